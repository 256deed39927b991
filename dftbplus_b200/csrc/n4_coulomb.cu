// n4_coulomb.cu -- the Coulomb 1/R matrix of the SCC shifts and its product with the charges on the
// device (second piece after Efilling):
//   invRClusterSerial   dftb/coulomb.F90:704-733     invRMat(i2, i1) = 1 / |r_i1 - r_i2|, i2 > i1
//   invRPeriodicSerial  dftb/coulomb.F90:968-1024    Ewald sum: real-space part over the dynamic neighbour
//                       list (addNeighbourContribsIS :1027-1053, type/dynneighlist.F90:207-260),
//                       reciprocal part ewaldReciprocal (:2079-2117), -pi / (V alpha^2), -2 alpha / sqrt(pi)
//   updateShifts        dftb/coulomb.F90:537-586     shiftPerAtom = hemv(invRMat, deltaQAtom, 'L')
// The reference does the Ewald matrix once per geometry on the host: nAtom^2 / 2 entries, each a loop over
// all lattice points inside the real-space cutoff plus nG cosines -- 4 * 10^9 cosines at 4096 atoms.
//
// Reciprocal part as a tensor-core contraction: cos(G.(r1 - r2)) = cos(G.r1) cos(G.r2) + sin(G.r1) sin(G.r2),
// so the reciprocal matrix is F diag(w) F^T with the structure factors F = [cos(G.r_i) | sin(G.r_i)]
// (nAtom x 2 nG) and w_G = 8 pi / V exp(-G^2 / 4 alpha^2) / G^2 -- exactly the weighted Gram matrix of K3
// (k3_density_matrix.cu: DMMA, TMA-fed): nAtom^2 nG FMAs on the FP64 tensor pipe instead of nAtom^2 nG / 2
// cosines on the CUDA cores.  The sum differs from the reference's term-by-term cosine sum by rounding
// only (a few 1e-16 of sum w_G).
// Real-space part: one thread per entry (i2 >= i1) walks the lattice translations in the reference's
// order (central cell first, then the iterator's order), so the sum of every entry has the reference's
// order of terms; -pi / (V alpha^2) and the self term are added in the same pass.
// Shifts: one pass over the lower triangle (the symmetric matrix-vector product touches every stored
// element once), row and column contributions through warp reductions + atomics.
#include "common.cuh"

using namespace dftbp_b200;

namespace dftbp_b200 {
int weightedGramLower(dftbp_b200_ctx *ctx, const double *dC, int ldc, const double *dW, int n, int K, double *dDm,
                      int ldd);
}

namespace {

constexpr double kPi = 3.14159265358979323846264338327950288;

__global__ void __launch_bounds__(256) n4_invr_cluster(int nAtom, const double *__restrict__ coord,
                                                       double *__restrict__ invR, int ld) {
  const long long total = (long long)nAtom * nAtom;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int i2 = (int)(t % nAtom), i1 = (int)(t / nAtom);
    double v = 0.0;
    if (i2 > i1) {
      const double x = coord[3 * i1] - coord[3 * i2], y = coord[3 * i1 + 1] - coord[3 * i2 + 1],
                   z = coord[3 * i1 + 2] - coord[3 * i2 + 2];
      v = 1.0 / sqrt(__dadd_rn(__dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y)), __dmul_rn(z, z)));
    }
    invR[i2 + (size_t)ld * i1] = v;
  }
}

// F(i, g) = cos(G_g . r_i), F(i, nG + g) = sin(G_g . r_i); w(g) = w(nG + g) = 8 pi / V exp(-G^2 / (4 alpha^2)) / G^2
__global__ void __launch_bounds__(256) n4_structure_factors(int nAtom, int nG, const double *__restrict__ coord,
                                                            const double *__restrict__ gLat, double alpha,
                                                            double volume, double *__restrict__ F, int ldf,
                                                            double *__restrict__ w) {
  const long long total = (long long)nAtom * nG;
  const double g2Factor = -1.0 / (4.0 * alpha * alpha);
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(t % nAtom), g = (int)(t / nAtom);
    const double gx = gLat[3 * g], gy = gLat[3 * g + 1], gz = gLat[3 * g + 2];
    double s, c;
    sincos(gx * coord[3 * i] + gy * coord[3 * i + 1] + gz * coord[3 * i + 2], &s, &c);
    F[i + (size_t)ldf * g] = c;
    F[i + (size_t)ldf * (nG + g)] = s;
    if (i == 0) {
      const double g2 = gx * gx + gy * gy + gz * gz;
      const double wg = 2.0 * (exp(g2Factor * g2) / g2) * 4.0 * kPi / volume;
      w[g] = wg;
      w[nG + g] = wg;
    }
  }
}

// invR(i2, i1) += sum over the lattice translations of erfc(alpha r) / r - pi / (V alpha^2) [- 2 alpha / sqrt(pi)]
__global__ void __launch_bounds__(256) n4_ewald_real(int nAtom, const double *__restrict__ coord, int nRLat,
                                                     const double *__restrict__ rLat, double alpha, double volume,
                                                     double cutoff2, double *__restrict__ invR, int ld) {
  const long long total = (long long)nAtom * nAtom;
  const double background = kPi / (volume * alpha * alpha), self = 2.0 * alpha / sqrt(kPi);
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int i2 = (int)(t % nAtom), i1 = (int)(t / nAtom);
    if (i2 < i1) continue;
    const double x1 = coord[3 * i1], y1 = coord[3 * i1 + 1], z1 = coord[3 * i1 + 2];
    const double x2 = coord[3 * i2], y2 = coord[3 * i2 + 1], z2 = coord[3 * i2 + 2];
    double sum = 0.0;
    for (int iL = (i2 == i1 ? 1 : 0); iL < nRLat; ++iL) {
      // neighCoords = coords0(:, iAtom2) + cellVec; dist2 = sum((coordsAtom1 - neighCoords)**2)
      const double dx = x1 - __dadd_rn(x2, __ldg(rLat + 3 * iL)), dy = y1 - __dadd_rn(y2, __ldg(rLat + 3 * iL + 1)),
                   dz = z1 - __dadd_rn(z2, __ldg(rLat + 3 * iL + 2));
      const double dist2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
      if (dist2 <= cutoff2) {
        const double rr = sqrt(dist2);
        sum += erfc(alpha * rr) / rr;
      }
    }
    double *p = invR + i2 + (size_t)ld * i1;
    double v = (sum + *p) - background;  // (real part + reciprocal part) - pi / (V alpha^2), the reference's order
    if (i2 == i1) v -= self;
    *p = v;
  }
}

// y += A x for the symmetric A stored in its lower triangle: tiles of 64 x 64, 256 threads (thread = row r of
// the tile, a group of 16 columns); y must be zero on entry
constexpr int kSymvTile = 64;
__global__ void __launch_bounds__(256) n4_symv_lower(int n, const double *__restrict__ A, int ld,
                                                     const double *__restrict__ x, double *__restrict__ y) {
  const int nT = (n + kSymvTile - 1) / kSymvTile;
  const long long nTiles = (long long)nT * (nT + 1) / 2;
  const int r = threadIdx.x & 63, cg = threadIdx.x >> 6, lane = threadIdx.x & 31;
  for (long long tile = blockIdx.x; tile < nTiles; tile += gridDim.x) {
    // tile -> (ti >= tj), column-major over the lower triangle of tiles
    int tj = 0;
    long long rest = tile;
    while (rest >= nT - tj) { rest -= nT - tj; ++tj; }
    const int ti = tj + (int)rest;
    const int i = ti * kSymvTile + r;
    const double xi = i < n ? __ldg(x + i) : 0.0;
    double rowSum = 0.0;
#pragma unroll 4
    for (int c = 0; c < 16; ++c) {
      const int j = tj * kSymvTile + cg * 16 + c;
      double a = 0.0;
      if (i < n && j < n && i >= j) a = __ldg(A + i + (size_t)ld * j);
      rowSum = fma(a, __ldg(x + min(j, n - 1)), rowSum);
      // column contribution of the strictly lower elements: y(j) += a x(i)
      double colPart = i > j ? a * xi : 0.0;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) colPart += __shfl_xor_sync(0xffffffffu, colPart, o);
      if (lane == 0 && j < n && colPart != 0.0) atomicAdd(y + j, colPart);
    }
    if (i < n && rowSum != 0.0) atomicAdd(y + i, rowSum);
  }
}

int gridOf(dftbp_b200_ctx *ctx, long long items, int block) {
  const long long want = (items + block - 1) / block, cap = (long long)ctx->smCount * 16;
  return (int)std::max<long long>(1, std::min(want, cap));
}

}  // namespace

extern "C" {

int dftbp_b200_invr_cluster(dftbp_b200_ctx *ctx, int nAtom, const double *coord, double *invRMat, int ld) {
  if (!ctx) return DFTBP_B200_EINVAL;
  DFTBP_CHECK(ctx, nAtom > 0 && coord && invRMat && ld >= nAtom, "invr_cluster: bad arguments");
  const double *dCoord;
  double *dA;
  DFTBP_TRY(devIn(ctx, kSlotA, coord, (size_t)3 * nAtom, &dCoord));
  DFTBP_TRY(devOut(ctx, kSlotB, invRMat, (size_t)ld * nAtom, false, &dA));
  if (ld > nAtom) DFTBP_CUDA(ctx, cudaMemsetAsync(dA, 0, sizeof(double) * (size_t)ld * nAtom, ctx->stream));
  n4_invr_cluster<<<gridOf(ctx, (long long)nAtom * nAtom, 256), 256, 0, ctx->stream>>>(nAtom, dCoord, dA, ld);
  ctx->launches++;
  return finish(ctx);
}

int dftbp_b200_invr_periodic(dftbp_b200_ctx *ctx, int nAtom, const double *coord, int nRLat, const double *rLat,
                             int nG, const double *gLat, double alpha, double volume, double rCutoff,
                             double *invRMat, int ld) {
  if (!ctx) return DFTBP_B200_EINVAL;
  DFTBP_CHECK(ctx, nAtom > 0 && coord && invRMat && ld >= nAtom, "invr_periodic: bad arguments");
  DFTBP_CHECK(ctx, nRLat >= 1 && rLat && nG >= 1 && gLat, "invr_periodic: lattice points missing");
  DFTBP_CHECK(ctx, alpha > 0.0 && volume > 0.0 && rCutoff >= 0.0, "invr_periodic: bad Ewald parameters");
  const double *dCoord, *dR, *dG;
  double *dA;
  DFTBP_TRY(devIn(ctx, kSlotA, coord, (size_t)3 * nAtom, &dCoord));
  DFTBP_TRY(devIn(ctx, kSlotC, rLat, (size_t)3 * nRLat, &dR));
  DFTBP_TRY(devIn(ctx, kSlotD, gLat, (size_t)3 * nG, &dG));
  DFTBP_TRY(devOut(ctx, kSlotB, invRMat, (size_t)ld * nAtom, false, &dA));
  // structure factors F(ldf, 2 nG) and weights w(2 nG) in one scratch block; ldf even for the bulk copies
  const int ldf = (nAtom + 1) & ~1;
  void *scratch;
  DFTBP_TRY(stagingGet(ctx, kSlotE, sizeof(double) * ((size_t)ldf * 2 * nG + 2 * (size_t)nG), &scratch));
  double *dF = static_cast<double *>(scratch), *dW = dF + (size_t)ldf * 2 * nG;
  if (ldf > nAtom) DFTBP_CUDA(ctx, cudaMemsetAsync(dF, 0, sizeof(double) * (size_t)ldf * 2 * nG, ctx->stream));
  n4_structure_factors<<<gridOf(ctx, (long long)nAtom * nG, 256), 256, 0, ctx->stream>>>(nAtom, nG, dCoord, dG, alpha,
                                                                                        volume, dF, ldf, dW);
  ctx->launches++;
  // reciprocal part: invR = lower(F diag(w) F^T), everything else zero (invRMat(:,:) = 0)
  DFTBP_TRY(weightedGramLower(ctx, dF, ldf, dW, nAtom, 2 * nG, dA, ld));
  n4_ewald_real<<<gridOf(ctx, (long long)nAtom * nAtom, 256), 256, 0, ctx->stream>>>(nAtom, dCoord, nRLat, dR, alpha, volume,
                                                                                    rCutoff * rCutoff, dA, ld);
  ctx->launches++;
  return finish(ctx);
}

int dftbp_b200_coulomb_shift(dftbp_b200_ctx *ctx, int nAtom, const double *invRMat, int ld, const double *deltaQAtom,
                             double *shiftPerAtom) {
  if (!ctx) return DFTBP_B200_EINVAL;
  DFTBP_CHECK(ctx, nAtom > 0 && invRMat && deltaQAtom && shiftPerAtom && ld >= nAtom, "coulomb_shift: bad arguments");
  const double *dA, *dQ;
  double *dY;
  DFTBP_TRY(devIn(ctx, kSlotA, invRMat, (size_t)ld * nAtom, &dA));
  DFTBP_TRY(devIn(ctx, kSlotC, deltaQAtom, (size_t)nAtom, &dQ));
  DFTBP_TRY(devOut(ctx, kSlotD, shiftPerAtom, (size_t)nAtom, false, &dY));
  DFTBP_CUDA(ctx, cudaMemsetAsync(dY, 0, sizeof(double) * nAtom, ctx->stream));  // this%shiftPerAtom_(:) = 0.0_dp
  const int nT = (nAtom + kSymvTile - 1) / kSymvTile;
  const long long nTiles = (long long)nT * (nT + 1) / 2;
  n4_symv_lower<<<(int)std::min<long long>(nTiles, (long long)ctx->smCount * 8), 256, 0, ctx->stream>>>(nAtom, dA, ld, dQ, dY);
  ctx->launches++;
  return finish(ctx);
}

}  // extern "C"

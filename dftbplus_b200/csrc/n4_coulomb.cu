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

// ---- short-range gamma (dftb/shortgamma.F90, dftb/shortgammafuncs.F90) ---------------------------------------
constexpr double kMinHubDiff = 0.3125e-5, kTolSameDist = 1.0e-5, kTolSameDist2 = 1.0e-5 * 1.0e-5;
constexpr int kMaxHubbU = 4;  // unique Hubbard U values per species (one per shell at most: s, p, d, f)

// gammaSubExprn_, shortgammafuncs.F90:470-500
__device__ __forceinline__ double gammaSubExprn(double rab, double tau1, double tau2) {
  const double d2 = tau1 * tau1 - tau2 * tau2, t24 = tau2 * tau2 * tau2 * tau2;
  return exp(-tau1 * rab) * ((0.5 * t24 * tau1 / (d2 * d2)) - (t24 * tau2 * tau2 - 3.0 * t24 * tau1 * tau1) / (rab * (d2 * d2 * d2)));
}

// expGamma, shortgammafuncs.F90:97-148
__device__ __forceinline__ double expGamma(double rab, double Ua, double Ub) {
  const double tauA = 3.2 * Ua, tauB = 3.2 * Ub;
  if (rab < kTolSameDist) {
    if (fabs(Ua - Ub) < kMinHubDiff) return -0.5 * (Ua + Ub);
    const double ab = tauA * tauB, sm = tauA + tauB;
    return -0.5 * (ab / sm + ab * ab / (sm * sm * sm));
  }
  if (fabs(Ua - Ub) < kMinHubDiff) {
    const double tauMean = 0.5 * (tauA + tauB);
    return exp(-tauMean * rab) * (1.0 / rab + 0.6875 * tauMean + 0.1875 * rab * (tauMean * tauMean) +
                                  0.02083333333333333333 * (rab * rab) * (tauMean * tauMean * tauMean));
  }
  return gammaSubExprn(rab, tauA, tauB) + gammaSubExprn(rab, tauB, tauA);
}

// search_asc_real_geq (math/binarysearch.F90:63-94) as getNrOfNeighbours calls it: xx1[jj] = xx(jj)
__device__ __forceinline__ int searchAscRealGeq(int n, const double *__restrict__ xx1, double xVal, double tol) {
  int jlower = 0, jupper = n;
  while (jlower < jupper) {
    const int jcurr = jlower + (jupper - jlower + 1) / 2;
    if (__dsub_rn(xVal, xx1[jcurr]) >= -tol) jlower = jcurr;
    else jupper = jcurr - 1;
  }
  return jlower;
}

struct GammaParams {
  int nAtom, ld, mU, nSpecies, mShell;
  const int *nHubbU;           // (nSpecies)
  const double *uniqHubbU;     // (mU, nSpecies)
  const double *shortCutoffs;  // (mU, mU, nSpecies, nSpecies)
  const int *species, *img2CentCell, *nNeighbour, *iNeighbour;
  const double *coords, *neighDist2;
  int *nNeighU;       // (mU, mU, nSpecies, nAtom)
  int *maxNeigh;      // (nAtom)
  double *shortGamma; // (mU, mU, 0:mNeigh, nAtom)
};

// updateNrOfNeighbours_ (shortgamma.F90:779-801): one thread per atom
__global__ void __launch_bounds__(128) n4_gamma_counts(const GammaParams p) {
  for (int iAt = blockIdx.x * blockDim.x + threadIdx.x; iAt < p.nAtom; iAt += gridDim.x * blockDim.x) {
    const int sp1 = p.species[iAt] - 1;
    int mx = 0;
    for (int iSp = 0; iSp < p.nSpecies; ++iSp)
      for (int iU1 = 0; iU1 < p.mU; ++iU1)
        for (int iU2 = 0; iU2 < p.mU; ++iU2) {
          int v = 0;
          if (iU1 < p.nHubbU[sp1] && iU2 < p.nHubbU[iSp]) {
            const double c = p.shortCutoffs[iU2 + p.mU * (iU1 + p.mU * (iSp + p.nSpecies * sp1))];
            v = searchAscRealGeq(p.nNeighbour[iAt], p.neighDist2 + (size_t)p.ld * iAt, c * c, kTolSameDist2);
          }
          p.nNeighU[iU2 + p.mU * (iU1 + p.mU * (iSp + p.nSpecies * iAt))] = v;
          mx = max(mx, v);
        }
    p.maxNeigh[iAt] = mx;
  }
}

// updateShortGammaValues_ (shortgamma.F90:804-857, no damping / H5): one thread per (iNeigh, iAtom) entry;
// shortGamma is zero on entry
__global__ void __launch_bounds__(256) n4_gamma_values(const GammaParams p) {
  const long long total = (long long)p.ld * p.nAtom;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int iAt = (int)(t / p.ld), iNeigh = (int)(t - (long long)iAt * p.ld);
    if (iNeigh > p.maxNeigh[iAt]) continue;
    const int iAt2 = p.iNeighbour[t] - 1, sp1 = p.species[iAt] - 1, sp2 = p.species[iAt2] - 1;
    const double dx = p.coords[3 * iAt] - p.coords[3 * iAt2], dy = p.coords[3 * iAt + 1] - p.coords[3 * iAt2 + 1],
                 dz = p.coords[3 * iAt + 2] - p.coords[3 * iAt2 + 2];
    const double rab = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz)));
    for (int iU1 = 0; iU1 < p.nHubbU[sp1]; ++iU1)
      for (int iU2 = 0; iU2 < p.nHubbU[sp2]; ++iU2)
        if (iNeigh <= p.nNeighU[iU2 + p.mU * (iU1 + p.mU * (sp2 + p.nSpecies * iAt))])
          p.shortGamma[iU2 + p.mU * (iU1 + p.mU * t)] = expGamma(rab, p.uniqHubbU[iU2 + p.mU * sp2], p.uniqHubbU[iU1 + p.mU * sp1]);
  }
}

// buildShifts_ (shortgamma.F90:860-933): one thread per (iNeigh, iAtom1) entry; shiftShell is zero on entry
__global__ void __launch_bounds__(256) n4_gamma_shifts(const GammaParams p, const int *__restrict__ nShell,
                                                       const int *__restrict__ iHubbU, const double *__restrict__ dq,
                                                       double *__restrict__ shiftShell) {
  const long long total = (long long)p.ld * p.nAtom;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int iAt1 = (int)(t / p.ld), iNeigh = (int)(t - (long long)iAt1 * p.ld);
    if (iNeigh > p.maxNeigh[iAt1]) continue;
    const int iAt2f = p.img2CentCell[p.iNeighbour[t] - 1] - 1, sp1 = p.species[iAt1] - 1, sp2 = p.species[iAt2f] - 1;
    const double *__restrict__ g = p.shortGamma + (size_t)p.mU * p.mU * t;
    for (int iSh1 = 0; iSh1 < nShell[sp1]; ++iSh1) {
      const int iU1 = iHubbU[iSh1 + p.mShell * sp1] - 1;
      const double dq1 = dq[iSh1 + p.mShell * iAt1];
      double acc = 0.0;
      for (int iSh2 = 0; iSh2 < nShell[sp2]; ++iSh2) {
        const double gv = g[(iHubbU[iSh2 + p.mShell * sp2] - 1) + p.mU * iU1];
        acc = fma(gv, dq[iSh2 + p.mShell * iAt2f], acc);
        if (iAt2f != iAt1 && gv != 0.0) atomicAdd(shiftShell + iSh2 + p.mShell * iAt2f, -gv * dq1);
      }
      if (acc != 0.0) atomicAdd(shiftShell + iSh1 + p.mShell * iAt1, -acc);
    }
  }
}

struct ShortGammaState {
  GammaParams p{};
  int *dNHubbU = nullptr, *dNNeighU = nullptr, *dMaxNeigh = nullptr;
  double *dUniq = nullptr, *dCut = nullptr, *dGamma = nullptr;
  size_t capNeighU = 0, capMax = 0, capGamma = 0, capU = 0, capCut = 0, capNH = 0;
  bool valid = false;
  unsigned long long generation = 0;  // ctx->listGeneration of the list the values belong to
};

template <class T>
int growTo(dftbp_b200_ctx *ctx, T **p, size_t *cap, size_t need) {
  if (*cap >= need) return 0;
  if (*p) cudaFree(*p);
  *p = nullptr;
  *cap = 0;
  cudaError_t e = cudaMalloc((void **)p, sizeof(T) * need);
  if (e != cudaSuccess) return fail(ctx, DFTBP_B200_ENOMEM, "short gamma: cudaMalloc failed: %s", cudaGetErrorString(e));
  *cap = need;
  return 0;
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


int dftbp_b200_short_gamma_update(dftbp_b200_ctx *ctx, int nSpecies, int mHubbU, const int *nHubbU,
                                  const double *uniqHubbU, const double *shortCutoffs) {
  if (!ctx) return DFTBP_B200_EINVAL;
  DFTBP_CHECK(ctx, nSpecies == ctx->nSpecies && nSpecies > 0, "short_gamma_update: species count differs from set_orbitals");
  DFTBP_CHECK(ctx, mHubbU >= 1 && mHubbU <= kMaxHubbU && nHubbU && uniqHubbU && shortCutoffs, "short_gamma_update: bad arguments");
  const NeighbourListView L = neighbourListView(ctx);
  DFTBP_CHECK(ctx, L.valid, "short_gamma_update: no neighbour list (call update_neighbour_list first)");
  if (!ctx->shortGammaState) ctx->shortGammaState = new ShortGammaState();
  ShortGammaState *st = static_cast<ShortGammaState *>(ctx->shortGammaState);
  st->valid = false;
  const int ld = L.mNeigh + 1;
  const size_t mu2 = (size_t)mHubbU * mHubbU;
  DFTBP_TRY(growTo(ctx, &st->dNHubbU, &st->capNH, (size_t)nSpecies));
  DFTBP_TRY(growTo(ctx, &st->dUniq, &st->capU, (size_t)mHubbU * nSpecies));
  DFTBP_TRY(growTo(ctx, &st->dCut, &st->capCut, mu2 * nSpecies * nSpecies));
  DFTBP_TRY(growTo(ctx, &st->dNNeighU, &st->capNeighU, mu2 * nSpecies * L.nAtom));
  DFTBP_TRY(growTo(ctx, &st->dMaxNeigh, &st->capMax, (size_t)L.nAtom));
  DFTBP_TRY(growTo(ctx, &st->dGamma, &st->capGamma, mu2 * ld * L.nAtom));
  auto put = [&](void *dst, const void *src, size_t bytes) -> int {
    DFTBP_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, isDevicePtr(src) ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice,
                                    ctx->stream));
    return 0;
  };
  DFTBP_TRY(put(st->dNHubbU, nHubbU, sizeof(int) * nSpecies));
  DFTBP_TRY(put(st->dUniq, uniqHubbU, sizeof(double) * mHubbU * nSpecies));
  DFTBP_TRY(put(st->dCut, shortCutoffs, sizeof(double) * mu2 * nSpecies * nSpecies));
  GammaParams &p = st->p;
  p.nAtom = L.nAtom; p.ld = ld; p.mU = mHubbU; p.nSpecies = nSpecies; p.mShell = ctx->mShell;
  p.nHubbU = st->dNHubbU; p.uniqHubbU = st->dUniq; p.shortCutoffs = st->dCut;
  p.species = L.species; p.img2CentCell = L.img2CentCell; p.nNeighbour = L.nNeighbour; p.iNeighbour = L.iNeighbour;
  p.coords = L.coords; p.neighDist2 = L.neighDist2;
  p.nNeighU = st->dNNeighU; p.maxNeigh = st->dMaxNeigh; p.shortGamma = st->dGamma;
  DFTBP_CUDA(ctx, cudaMemsetAsync(st->dGamma, 0, sizeof(double) * mu2 * ld * L.nAtom, ctx->stream));
  n4_gamma_counts<<<gridOf(ctx, L.nAtom, 128), 128, 0, ctx->stream>>>(p);
  n4_gamma_values<<<gridOf(ctx, (long long)ld * L.nAtom, 256), 256, 0, ctx->stream>>>(p);
  ctx->launches += 2;
  DFTBP_CUDA(ctx, cudaGetLastError());
  // host inputs were staged with asynchronous copies from pageable memory: consumed when the stream is idle
  DFTBP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  st->valid = true;
  st->generation = ctx->listGeneration;
  return 0;
}

int dftbp_b200_short_gamma_get(dftbp_b200_ctx *ctx, int *nNeighU, int *maxNeigh, double *shortGamma) {
  if (!ctx) return DFTBP_B200_EINVAL;
  ShortGammaState *st = static_cast<ShortGammaState *>(ctx->shortGammaState);
  DFTBP_CHECK(ctx, st && st->valid, "short_gamma_get: call short_gamma_update first");
  const GammaParams &p = st->p;
  const size_t mu2 = (size_t)p.mU * p.mU;
  auto get = [&](void *dst, const void *src, size_t bytes) -> int {
    if (!dst) return 0;
    DFTBP_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, isDevicePtr(dst) ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost,
                                    ctx->stream));
    return 0;
  };
  DFTBP_TRY(get(nNeighU, p.nNeighU, sizeof(int) * mu2 * p.nSpecies * p.nAtom));
  DFTBP_TRY(get(maxNeigh, p.maxNeigh, sizeof(int) * p.nAtom));
  DFTBP_TRY(get(shortGamma, p.shortGamma, sizeof(double) * mu2 * p.ld * p.nAtom));
  DFTBP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return 0;
}

int dftbp_b200_short_gamma_shifts(dftbp_b200_ctx *ctx, const int *iHubbU, const double *deltaQShell,
                                  double *shiftShell) {
  if (!ctx) return DFTBP_B200_EINVAL;
  ShortGammaState *st = static_cast<ShortGammaState *>(ctx->shortGammaState);
  DFTBP_CHECK(ctx, st && st->valid, "short_gamma_shifts: call short_gamma_update first");
  DFTBP_CHECK(ctx, iHubbU && deltaQShell && shiftShell, "short_gamma_shifts: null array");
  const GammaParams &p = st->p;
  {
    // the values belong to one neighbour list (TShortGamma%updateCoords follows every list update)
    const NeighbourListView L = neighbourListView(ctx);
    DFTBP_CHECK(ctx, L.valid && st->generation == ctx->listGeneration && L.iNeighbour == p.iNeighbour && L.nAtom == p.nAtom && L.mNeigh + 1 == p.ld,
                "short_gamma_shifts: the neighbour list changed since short_gamma_update");
  }
  const int *dIU, *dNShell;
  const double *dQ;
  double *dS;
  DFTBP_TRY(devIn(ctx, kSlotA, iHubbU, (size_t)p.mShell * p.nSpecies, &dIU));
  DFTBP_TRY(devIn(ctx, kSlotC, ctx->nShell.data(), (size_t)p.nSpecies, &dNShell));
  DFTBP_TRY(devIn(ctx, kSlotD, deltaQShell, (size_t)p.mShell * p.nAtom, &dQ));
  DFTBP_TRY(devOut(ctx, kSlotB, shiftShell, (size_t)p.mShell * p.nAtom, false, &dS));
  DFTBP_CUDA(ctx, cudaMemsetAsync(dS, 0, sizeof(double) * p.mShell * p.nAtom, ctx->stream));  // shiftShell(:,:) = 0.0_dp
  n4_gamma_shifts<<<gridOf(ctx, (long long)p.ld * p.nAtom, 256), 256, 0, ctx->stream>>>(p, dNShell, dIU, dQ, dS);
  ctx->launches++;
  // ctx->nShell is a host vector staged asynchronously: wait even when the outputs stay on the device
  DFTBP_TRY(finish(ctx));
  DFTBP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return 0;
}

}  // extern "C"

namespace dftbp_b200 {
void freeShortGamma(dftbp_b200_ctx *ctx) {
  ShortGammaState *st = static_cast<ShortGammaState *>(ctx->shortGammaState);
  if (!st) return;
  cudaFree(st->dNHubbU); cudaFree(st->dNNeighU); cudaFree(st->dMaxNeigh);
  cudaFree(st->dUniq); cudaFree(st->dCut); cudaFree(st->dGamma);
  delete st;
  ctx->shortGammaState = nullptr;
}
}  // namespace dftbp_b200

/* chain.c -- plain C99 driver of the hot path through include/dftbp_b200.h (no CUDA, no Python).
 *
 * An H2O-like molecule (species 1: s p, species 2: s) with smooth synthetic Slater-Koster tables
 * goes through the calls a DFTB+ maintainer would make per geometry and SCC iteration
 * (dftbplus/main.F90:1219-1221, 3536-3566, 4101-4108, 4705, 7264):
 *   set_orbitals -> set_sk_table -> set_geometry -> build_h0_s -> unpack_hs_real (H, S) ->
 *   hegv_real -> density_matrix_real -> pack_hs_real -> mulliken_orbital -> derivative_shift
 * with host buffers, and checks what must hold whatever the tables are: unit diagonal of S, the
 * electron count from the Mulliken charges, vanishing net force.  Exit code 0 = all checks passed,
 * 77 = no CUDA device (the library has no CPU fallback), anything else = failure.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "dftbp_b200.h"

#define CHECK(call)                                                                      \
  do {                                                                                   \
    int rc_ = (call);                                                                    \
    if (rc_ != 0) {                                                                      \
      fprintf(stderr, "%s -> %d: %s\n", #call, rc_, dftbp_b200_last_error(ctx));          \
      return 1;                                                                          \
    }                                                                                    \
  } while (0)

enum { N_GRID = 500, N_ATOM = 3, M_NEIGH = 2, N_ORB = 6, SPARSE = 27 };
static const double kDist = 0.02;

/* integral k of a table: smooth, decaying, distinct per (which, pair, k) */
static void fill_table(double *tab, int nInteg, int which, int seed) {
  for (int k = 0; k < nInteg; ++k)
    for (int j = 0; j < N_GRID; ++j) {
      const double r = (j + 1) * kDist;
      const double a = 0.9 + 0.13 * k + 0.07 * seed, s = (k & 1) ? -1.0 : 1.0;
      tab[j + (size_t)N_GRID * k] = (which ? 1.0 : -0.35) * s * (1.0 + 0.3 * r) * exp(-a * r) * (0.6 + 0.1 * k);
    }
}

int main(void) {
  dftbp_b200_ctx *ctx = NULL;
  int rc = dftbp_b200_create(&ctx, 0, 0, 1);
  if (rc == DFTBP_B200_ENODEVICE) {
    printf("no CUDA device: %s\n", dftbp_b200_last_error(ctx));
    return 77;
  }
  if (rc != 0) return 1;
  printf("library %s\n", dftbp_b200_version());

  /* TOrbitals: species 1 = (s, p), species 2 = (s); angShell(mShell, nSpecies) */
  const int nShell[2] = {2, 1}, angShell[4] = {0, 1, 0, 0};
  CHECK(dftbp_b200_set_orbitals(ctx, 2, 2, nShell, angShell));
  /* integrals per ordered species pair: sum over shell pairs of min(l1, l2) + 1 */
  const int nInteg[2][2] = {{5, 2}, {2, 1}};
  double *tab = (double *)malloc(sizeof(double) * N_GRID * 5);
  for (int which = 0; which < 2; ++which)
    for (int sp1 = 1; sp1 <= 2; ++sp1)
      for (int sp2 = 1; sp2 <= 2; ++sp2) {
        /* the (sp1, sp2) and (sp2, sp1) tables of a real .skf pair are consistent; here the on-site
         * symmetric choice seed = sp1 + sp2 keeps S symmetric enough to be positive definite */
        fill_table(tab, nInteg[sp1 - 1][sp2 - 1], which, sp1 + sp2);
        CHECK(dftbp_b200_set_sk_table(ctx, which, sp1, sp2, N_GRID, nInteg[sp1 - 1][sp2 - 1], kDist, tab,
                                      DFTBP_B200_SK_EQGRID_NEW));
      }
  free(tab);
  double cutoff = 0.0;
  CHECK(dftbp_b200_get_sk_cutoff(ctx, &cutoff));
  printf("SK cutoff %.3f bohr\n", cutoff);

  /* geometry: O at the origin, two H (bohr); cluster: one cell vector, every atom its own image */
  const double coords[9] = {0.0, 0.0, 0.0, 0.0, 1.43, 1.11, 0.0, -1.43, 1.11};
  const int species[3] = {1, 2, 2}, img2CentCell[3] = {1, 2, 3}, iCellVec[3] = {1, 1, 1};
  const double cellVec[3] = {0.0, 0.0, 0.0};
  const int nNeighbourSK[3] = {2, 1, 0};
  /* iNeighbour(0:mNeigh, nAtom), iPair(0:mNeigh, nAtom): blocks nOrb(neighbour) x nOrb(atom), column-major */
  const int iNeighbour[9] = {1, 2, 3, 2, 3, 0, 3, 0, 0};
  const int iPair[9] = {0, 16, 20, 24, 25, 0, 26, 0, 0};
  const int iAtomStart[4] = {1, 5, 6, 7};
  CHECK(dftbp_b200_set_geometry(ctx, N_ATOM, N_ATOM, M_NEIGH, coords, species, img2CentCell, iCellVec, 1, cellVec,
                                nNeighbourSK, iNeighbour, iPair, iAtomStart, SPARSE));

  /* on-site energies selfegy(mShell, nSpecies) */
  const double selfegy[4] = {-0.88, -0.33, -0.24, 0.0};
  double ham[SPARSE], over[SPARSE];
  CHECK(dftbp_b200_build_h0_s(ctx, selfegy, ham, over));
  int bad = 0;
  /* on-site blocks of S are unit matrices (nonscc.F90:166-173) */
  for (int a = 0; a < 4; ++a)
    for (int b = 0; b < 4; ++b)
      if (fabs(over[b + 4 * a] - (a == b ? 1.0 : 0.0)) > 1e-15) bad = 1;
  if (fabs(over[24] - 1.0) > 1e-15 || fabs(over[26] - 1.0) > 1e-15) bad = 1;
  if (bad) { fprintf(stderr, "on-site overlap blocks are not unit matrices\n"); return 2; }

  double H[N_ORB * N_ORB], S[N_ORB * N_ORB], eigen[N_ORB], dm[N_ORB * N_ORB];
  CHECK(dftbp_b200_unpack_hs_real(ctx, ham, H, N_ORB));
  CHECK(dftbp_b200_unpack_hs_real(ctx, over, S, N_ORB));
  CHECK(dftbp_b200_hegv_real(ctx, H, S, eigen, N_ORB, N_ORB, 'V'));
  printf("eigenvalues:");
  for (int i = 0; i < N_ORB; ++i) printf(" %.6f", eigen[i]);
  printf("\n");
  for (int i = 1; i < N_ORB; ++i)
    if (!(eigen[i] >= eigen[i - 1])) { fprintf(stderr, "eigenvalues not ascending\n"); return 3; }

  /* 8 electrons: four doubly occupied levels */
  const double filling[N_ORB] = {2.0, 2.0, 2.0, 2.0, 0.0, 0.0};
  CHECK(dftbp_b200_density_matrix_real(ctx, dm, H, N_ORB, N_ORB, filling, NULL));
  double rho[SPARSE], erho[SPARSE], weights[N_ORB];
  memset(rho, 0, sizeof rho);
  memset(erho, 0, sizeof erho);
  CHECK(dftbp_b200_pack_hs_real(ctx, rho, dm, N_ORB));
  for (int i = 0; i < N_ORB; ++i) weights[i] = filling[i] * eigen[i];
  CHECK(dftbp_b200_density_matrix_real(ctx, dm, H, N_ORB, N_ORB, filling, weights));
  CHECK(dftbp_b200_pack_hs_real(ctx, erho, dm, N_ORB));

  double qq[4 * N_ATOM], nEl = 0.0;  /* qq(mOrb, nAtom) accumulates */
  memset(qq, 0, sizeof qq);
  CHECK(dftbp_b200_mulliken_orbital(ctx, qq, over, rho));
  for (int i = 0; i < 4 * N_ATOM; ++i) nEl += qq[i];
  printf("Mulliken electron count %.12f\n", nEl);
  if (fabs(nEl - 8.0) > 1e-10) { fprintf(stderr, "electron count is not 8\n"); return 4; }

  double deriv[3 * N_ATOM], net[3] = {0.0, 0.0, 0.0}, big = 0.0;
  CHECK(dftbp_b200_derivative_shift(ctx, deriv, rho, 1, erho, NULL, DFTBP_B200_DIFF_FINITE, 0.0));
  for (int i = 0; i < N_ATOM; ++i)
    for (int c = 0; c < 3; ++c) {
      net[c] += deriv[c + 3 * i];
      if (fabs(deriv[c + 3 * i]) > big) big = fabs(deriv[c + 3 * i]);
    }
  printf("largest gradient component %.6e, net %.2e %.2e %.2e\n", big, net[0], net[1], net[2]);
  if (!(big > 1e-6) || fabs(net[0]) + fabs(net[1]) + fabs(net[2]) > 1e-10) {
    fprintf(stderr, "gradients vanish or do not add up to zero\n");
    return 5;
  }
  /* SCC pieces: 1/R matrix of the cluster, Coulomb shifts for the Mulliken charges, one step of
   * the Broyden mixer -- TCoulomb%updateCoords / %updateShifts, TBroydenMixerReal%mix from plain C */
  {
    double invR[N_ATOM * N_ATOM], dq[N_ATOM], shiftAtom[N_ATOM], qIn[N_ATOM], qDiff[N_ATOM];
    const double q0[N_ATOM] = {6.0, 1.0, 1.0};
    CHECK(dftbp_b200_invr_cluster(ctx, N_ATOM, coords, invR, N_ATOM));
    for (int j = 0; j < N_ATOM; ++j)
      for (int i = 0; i < N_ATOM; ++i) {
        const double dx = coords[3 * i] - coords[3 * j], dy = coords[3 * i + 1] - coords[3 * j + 1],
                     dz = coords[3 * i + 2] - coords[3 * j + 2];
        const double want = i > j ? 1.0 / sqrt(dx * dx + dy * dy + dz * dz) : 0.0;
        if (fabs(invR[i + N_ATOM * j] - want) > 1e-15) { fprintf(stderr, "invRCluster differs\n"); return 6; }
      }
    double sumDq = 0.0;
    for (int a = 0; a < N_ATOM; ++a) {
      dq[a] = qq[4 * a] + qq[4 * a + 1] + qq[4 * a + 2] + qq[4 * a + 3] - q0[a];
      sumDq += dq[a];
    }
    CHECK(dftbp_b200_coulomb_shift(ctx, N_ATOM, invR, N_ATOM, dq, shiftAtom));
    for (int a = 0; a < N_ATOM; ++a) {
      double want = 0.0;
      for (int b = 0; b < N_ATOM; ++b)
        if (b != a) want += invR[(a > b ? a : b) + N_ATOM * (a > b ? b : a)] * dq[b];
      if (fabs(shiftAtom[a] - want) > 1e-14) { fprintf(stderr, "Coulomb shift differs\n"); return 7; }
    }
    printf("charge transfer %.6f %.6f %.6f (sum %.1e), Coulomb shifts %.6f %.6f %.6f\n", dq[0], dq[1], dq[2], sumDq,
           shiftAtom[0], shiftAtom[1], shiftAtom[2]);
    CHECK(dftbp_b200_mixer_init(ctx, 100, 0.2, 0.01, 1.0, 1.0e5, 1.0e-2));
    CHECK(dftbp_b200_mixer_reset(ctx, N_ATOM));
    for (int a = 0; a < N_ATOM; ++a) { qIn[a] = q0[a]; qDiff[a] = dq[a]; }
    CHECK(dftbp_b200_mixer_mix(ctx, qIn, qDiff, N_ATOM));
    for (int a = 0; a < N_ATOM; ++a)   /* first iteration: simple mixing with mixParam (broydenmixer.F90:268-273) */
      if (fabs(qIn[a] - (q0[a] + 0.2 * dq[a])) > 1e-15) { fprintf(stderr, "first Broyden step is not simple mixing\n"); return 8; }
  }
  printf("kernel launches %lld, pairs %lld\n", dftbp_b200_kernel_launches(ctx), dftbp_b200_num_pairs(ctx));
  CHECK(dftbp_b200_destroy(ctx));
  printf("chain.c: OK\n");
  return 0;
}

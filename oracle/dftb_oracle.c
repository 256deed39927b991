/*
 * dftb_oracle.c -- CPU restatement of the DFTB+ per-SCC-step hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the parity oracle: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * build, load or call it.  Nothing under dftbplus_b200/ links or imports it.
 *
 * PARITY STATUS: not checked against a run of the Fortran BINARY -- this container has no Fortran compiler and no
 * .skf parameter files -- but pinned to the reference's own SOURCE, executed here by other means:
 *   - buildH0 / buildS / buildDiatomicBlocks / rotateH0, unpackHS / packHS (real and k-points), mullikenPerOrbital /
 *     PerBlock, addShift_atom / _block, derivativeNonSccEuclidian / derivative_blockEuclidean and getNonSCCStress /
 *     getBlockStress with getFirstDerivFiniteDiff (and, for oracle.py, fullDensityMatrix_real / fullEnergyDensityMatrix_real): the reference's source text is transpiled statement by statement to NumPy with Fortran
 *     index semantics and RUN (oracle/fortran_exec.py, oracle/gen_ref_sparse.py); its outputs are the golden vectors
 *     tests/golden/sparse_ref_golden.npz, which this file reproduces (H0 / S to 5e-17, forces to 2e-12, the rest to
 *     rounding; tests/test_oracle_sparse_ref.py); updateNeighbourList the same way (oracle/gen_ref_neigh.py,
 *     tests/test_oracle_neigh_ref.py): every array bit for bit, ties and the "too close" pair included; the SK-table
 *     interpolation SlakoEqGrid_interNew_ / _interOld_ (oracle/gen_ref_interp.py): bit for bit in every branch;
 *   - the shell-pair rotations ss ... ff, the interpolation routines, the tolerance heap sort and the scalar
 *     SCC-potential functions: machine translations to C (oracle/gen_ref_sk.py, gen_ref_math.py, gen_ref_scc.py ->
 *     oracle/_ref/, golden vectors under tests/golden/), bit for bit where the operation order is the reference's;
 *   - chunk ranges, the bisection search_asc_real_geq, the neighbour list's proximity rule and the Broyden mixer: the
 *     known answers of the reference's own unit tests (unit/common/schedule.F90, unit/math/binarysearch.F90,
 *     unit/dftb/periodic.F90, unit/mixer/broyden.F90).
 * (scc_model.py's Efilling: the same, oracle/gen_ref_filling.py; the Ewald / cluster 1/R matrix with the dynamic neighbour
 * list iterator: oracle/gen_ref_ewald.py; the short-range gamma: oracle/gen_ref_gamma.py.)  Still restated only: the
 * eigensolve (SciPy).
 *
 * Every routine cites the reference file:line (relative to /root/reference/src/dftbp/)
 * whose arithmetic and loop order it follows.  Arrays use the reference's layout:
 * Fortran column-major, 1-based atom / species values, 0-based iPair offsets.
 *
 * Build: gcc -O2 -funroll-all-loops -fopenmp -ffp-contract=off -fPIC -shared
 * (the flags of the reference's Release build, sys/gnu.cmake:27; x86-64 baseline => no FMA).
 */
#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_MAXL 3
#define ORC_MAXORB 16          /* s+p+d+f */
#define ORC_MAXINT 64          /* >= max nInteg per species pair */

/* common/accuracy.F90:44,48,69-82 */
static const double distFudge = 1.0;
static const double distFudgeOld = 0.3;
static const double tolSameDist2 = 1.0e-5 * 1.0e-5;
static const double minNeighDist2 = 1.0e-2 * 1.0e-2;

/* ------------------------------------------------------------------------------------------
 * math/interpolation.F90
 * ---------------------------------------------------------------------------------------- */

/* poly5ToZero, math/interpolation.F90:31-67 */
static double poly5ToZero(double y0, double y0p, double y0pp, double xx, double dx, double invdx)
{
  double dx1 = y0p * dx;
  double dx2 = y0pp * dx * dx;
  double dd = 10.0 * y0 - 4.0 * dx1 + 0.5 * dx2;
  double ee = -15.0 * y0 + 7.0 * dx1 - 1.0 * dx2;
  double ff = 6.0 * y0 - 3.0 * dx1 + 0.5 * dx2;
  double xr = xx * invdx;
  return ((ff * xr + ee) * xr + dd) * xr * xr * xr;
}

/* freeCubicSpline, math/interpolation.F90:74-120 (yy / yp+ypp both optional via NULL) */
static void freeCubicSpline(double y0, double y0p, double y0pp, double dx, double ydx, double xx,
                            double *yy, double *yp, double *ypp)
{
  double aa = y0, bb = y0p, cc = 0.5 * y0pp, dx1 = 1.0 / dx;
  double dd = (((ydx - y0) * dx1 - y0p) * dx1 - 0.5 * y0pp) * dx1;
  if (yy) *yy = ((dd * xx + cc) * xx + bb) * xx + aa;
  if (yp) {
    *yp = (3.0 * dd * xx + 2.0 * cc) * xx + bb;
    *ypp = 6.0 * dd * xx + 2.0 * cc;
  }
}

/* polyInterU2, math/interpolation.F90:182-239.  Neville tableau on a uniform grid, vectorised
 * over the integrals.  yp is (nInt, nn) column-major as in the reference (yp(:,ii)). */
static void polyInterU2(int nn, const double *xp, int nInt, const double *yp, double xx,
                        double *yy)
{
  double cc[ORC_MAXINT * 8], dd[ORC_MAXINT * 8], delta[8];
  int ii, mm, kk, iCl;
  for (ii = 1; ii <= nn - 1; ++ii) delta[ii - 1] = 1.0 / (xp[ii] - xp[0]);
  memcpy(cc, yp, sizeof(double) * nInt * nn);
  memcpy(dd, yp, sizeof(double) * nInt * nn);
  iCl = (int)ceil((xx - xp[0]) * delta[0]);
  /* the reference indexes yp(:,iCl) unguarded; distances below the first grid point
   * (r < dist) would read out of bounds there -- we clamp instead of invoking UB. */
  if (iCl < 1) iCl = 1;
  if (iCl > nn) iCl = nn;
  for (kk = 0; kk < nInt; ++kk) yy[kk] = yp[kk + nInt * (iCl - 1)];
  iCl -= 1;
  for (mm = 1; mm <= nn - 1; ++mm) {
    for (ii = 1; ii <= nn - mm; ++ii) {
      double x1 = xp[ii - 1] - xx, x2 = xp[ii + mm - 1] - xx;
      for (kk = 0; kk < nInt; ++kk) {
        double r2 = (dd[kk + nInt * (ii - 1)] - cc[kk + nInt * ii]) * delta[mm - 1];
        cc[kk + nInt * (ii - 1)] = x1 * r2;
        dd[kk + nInt * (ii - 1)] = x2 * r2;
      }
    }
    if (2 * iCl < nn - mm) {
      for (kk = 0; kk < nInt; ++kk) yy[kk] = yy[kk] + cc[kk + nInt * iCl];
    } else {
      for (kk = 0; kk < nInt; ++kk) yy[kk] = yy[kk] + dd[kk + nInt * (iCl - 1)];
      iCl -= 1;
    }
  }
}

/* exported for tests/test_oracle_math_ref.py: these restatements against the machine translation of the
 * reference's own routines (oracle/gen_ref_math.py -> oracle/_ref/libmathref.so) and its golden vectors */
double orc_poly5_to_zero(double y0, double y0p, double y0pp, double xx, double dx, double invdx)
{
  return poly5ToZero(y0, y0p, y0pp, xx, dx, invdx);
}

void orc_free_cubic_spline(double y0, double y0p, double y0pp, double dx, double ydx, double xx,
                           double *yy, double *yp, double *ypp)
{
  freeCubicSpline(y0, y0p, y0pp, dx, ydx, xx, yy, yp, ypp);
}

void orc_poly_inter_u2(int nn, const double *xp, int nInt, const double *yp, double xx, double *yy)
{
  polyInterU2(nn, xp, nInt, yp, xx, yy);
}

/* ------------------------------------------------------------------------------------------
 * dftb/slakoeqgrid.F90 : one equidistant SK table.  skTab(nGrid, nInteg) column-major.
 * ---------------------------------------------------------------------------------------- */
typedef struct {
  int nGrid, nInteg, method; /* method: 1 = skEqGridOld, 2 = skEqGridNew (slakoeqgrid.F90:62-65) */
  double dist;
  const double *tab;
} orc_table;

/* SlakoEqGrid_interNew_, slakoeqgrid.F90:183-244 */
static void interNew(const orc_table *t, double *dd, double rr)
{
  enum { nInter = 8, nRight = 4 };
  const double deltaR = 1e-5;
  double xa[nInter], yb[ORC_MAXINT * nInter];
  int leng = t->nGrid, nI = t->nInteg, ii, jj, iLast;
  double incr = t->dist;
  double rMax = (double)leng * incr + distFudge;
  int ind = (int)floor(rr / incr);
  for (ii = 0; ii < nI; ++ii) dd[ii] = 0.0;
  if (rr >= rMax) {
    return;
  } else if (ind < leng) {
    iLast = ind + nRight < leng ? ind + nRight : leng;
    if (iLast < nInter) iLast = nInter;
    for (ii = 1; ii <= nInter; ++ii) xa[ii - 1] = (double)(iLast - nInter + ii) * incr;
    for (jj = 0; jj < nI; ++jj)
      for (ii = 0; ii < nInter; ++ii)
        yb[jj + nI * ii] = t->tab[(iLast - nInter + ii) + (size_t)leng * jj];
    polyInterU2(nInter, xa, nI, yb, rr, dd);
  } else {
    double y0[ORC_MAXINT], y2[ORC_MAXINT];
    double dr = rr - rMax;
    iLast = leng;
    for (ii = 1; ii <= nInter; ++ii) xa[ii - 1] = (double)(iLast - nInter + ii) * incr;
    for (jj = 0; jj < nI; ++jj)
      for (ii = 0; ii < nInter; ++ii)
        yb[jj + nI * ii] = t->tab[(iLast - nInter + ii) + (size_t)leng * jj];
    polyInterU2(nInter, xa, nI, yb, xa[nInter - 1] - deltaR, y0);
    polyInterU2(nInter, xa, nI, yb, xa[nInter - 1] + deltaR, y2);
    for (ii = 0; ii < nI; ++ii) {
      double y1 = t->tab[(leng - 1) + (size_t)leng * ii];
      double y1p = (y2[ii] - y0[ii]) / (2.0 * deltaR);
      double y1pp = (y2[ii] + y0[ii] - 2.0 * y1) / (deltaR * deltaR);
      dd[ii] = poly5ToZero(y1, y1p, y1pp, dr, -1.0 * distFudge, -1.0 / distFudge);
    }
  }
}

/* SlakoEqGrid_interOld_, slakoeqgrid.F90:248-317 */
static void interOld(const orc_table *t, double *dd, double rr)
{
  enum { nInter = 3, nRight = 2 };
  double xa[nInter], yb[ORC_MAXINT * nInter];
  int leng = t->nGrid, nI = t->nInteg, ii, jj, iLast;
  double incr = t->dist;
  int mInd = leng + (int)floor(distFudgeOld / incr);
  int ind = (int)floor(rr / incr);
  double invdistFudge = -1.0 / ((double)(mInd - leng - 1) * incr);
  for (ii = 0; ii < nI; ++ii) dd[ii] = 0.0;
  if (ind < leng - 1) {
    iLast = ind + nRight < leng ? ind + nRight : leng;
    if (iLast < nInter) iLast = nInter;
    for (ii = 1; ii <= nInter; ++ii) xa[ii - 1] = (double)(iLast - nInter + ii) * incr;
    for (jj = 0; jj < nI; ++jj)
      for (ii = 0; ii < nInter; ++ii)
        yb[jj + nI * ii] = t->tab[(iLast - nInter + ii) + (size_t)leng * jj];
    polyInterU2(nInter, xa, nI, yb, rr, dd);
  } else if (ind < leng) {
    double dr = rr - (double)(leng - 1) * incr;
    for (ii = 0; ii < nI; ++ii) {
      const double *c = t->tab + (size_t)leng * ii;
      double y0 = c[leng - 3], y1 = c[leng - 2], y2 = c[leng - 1];
      double y1p = (y2 - y0) / (2.0 * incr);
      double y1pp = (y2 + y0 - 2.0 * y1) / (incr * incr);
      freeCubicSpline(y1, y1p, y1pp, incr, y2, dr, &dd[ii], NULL, NULL);
    }
  } else if (ind < mInd - 1) {
    double dr = rr - (double)(mInd - 1) * incr;
    for (ii = 0; ii < nI; ++ii) {
      const double *c = t->tab + (size_t)leng * ii;
      double y0 = c[leng - 3], y1 = c[leng - 2], y2 = c[leng - 1], y1p, y1pp;
      double r1 = (y2 - y0) / (2.0 * incr);
      double r2 = (y2 + y0 - 2.0 * y1) / (incr * incr);
      freeCubicSpline(y1, r1, r2, incr, y2, incr, NULL, &y1p, &y1pp);
      dd[ii] = poly5ToZero(y2, y1p, y1pp, dr, -1.0 * (double)(mInd - leng - 1) * incr,
                           invdistFudge);
    }
  }
}

/* SlakoEqGrid_getSKIntegrals, slakoeqgrid.F90:127-147 */
void orc_sk_eval(int nGrid, int nInteg, double dist, const double *tab, int method, double rr,
                 double *sk)
{
  orc_table t = {nGrid, nInteg, method, dist, tab};
  if (method == 1) interOld(&t, sk, rr); else interNew(&t, sk, rr);
}

/* SlakoEqGrid_getCutoff, slakoeqgrid.F90:162-179 */
double orc_sk_cutoff(int nGrid, double dist, int method)
{
  return (double)nGrid * dist + (method == 1 ? distFudgeOld : distFudge);
}

/* ------------------------------------------------------------------------------------------
 * Model description = flattened TOrbitals (type/orbitals.F90:20-51) + TSlakoCont
 * (dftb/slakocont.F90:32-40).  tabH/tabS[sp1 + nSpecies*sp2] (0-based species) is the table
 * the reference fetches with getSKIntegrals(skCont, sk, dist, sp1, sp2) (slakocont.F90:156-176),
 * i.e. sp1 = species of the atom, sp2 = species of its neighbour.
 * ---------------------------------------------------------------------------------------- */
typedef struct {
  int nSpecies, mShell, mOrb;
  const int *nShell;      /* (nSpecies) */
  const int *angShell;    /* (mShell, nSpecies) column-major */
  const int *nOrbSpecies; /* (nSpecies) */
  orc_table *tabH, *tabS; /* nSpecies*nSpecies each */
} orc_model;

orc_model *orc_model_create(int nSpecies, int mShell, const int *nShell, const int *angShell)
{
  orc_model *m = (orc_model *)calloc(1, sizeof(orc_model));
  int *ns = (int *)malloc(sizeof(int) * nSpecies);
  int *as = (int *)malloc(sizeof(int) * nSpecies * mShell);
  int *no = (int *)malloc(sizeof(int) * nSpecies);
  int iSp, iSh;
  memcpy(ns, nShell, sizeof(int) * nSpecies);
  memcpy(as, angShell, sizeof(int) * nSpecies * mShell);
  m->nSpecies = nSpecies; m->mShell = mShell; m->mOrb = 0;
  for (iSp = 0; iSp < nSpecies; ++iSp) {
    no[iSp] = 0;
    for (iSh = 0; iSh < ns[iSp]; ++iSh) no[iSp] += 2 * as[iSh + mShell * iSp] + 1;
    if (no[iSp] > m->mOrb) m->mOrb = no[iSp];
  }
  m->nShell = ns; m->angShell = as; m->nOrbSpecies = no;
  m->tabH = (orc_table *)calloc((size_t)nSpecies * nSpecies, sizeof(orc_table));
  m->tabS = (orc_table *)calloc((size_t)nSpecies * nSpecies, sizeof(orc_table));
  return m;
}

/* SlakoCont_addTableEqGrid, slakocont.F90:97-111 (the table is copied, as in SlakoEqGrid_init
 * slakoeqgrid.F90:95-121) */
void orc_model_set_table(orc_model *m, int which, int sp1, int sp2, int nGrid, int nInteg,
                         double dist, const double *tab, int method)
{
  orc_table *t = (which == 0 ? m->tabH : m->tabS) + (sp1 - 1) + m->nSpecies * (sp2 - 1);
  double *cp = (double *)malloc(sizeof(double) * nGrid * nInteg);
  memcpy(cp, tab, sizeof(double) * nGrid * nInteg);
  free((void *)t->tab);
  t->nGrid = nGrid; t->nInteg = nInteg; t->dist = dist; t->tab = cp; t->method = method;
}

void orc_model_destroy(orc_model *m)
{
  int i;
  if (!m) return;
  for (i = 0; i < m->nSpecies * m->nSpecies; ++i) { free((void *)m->tabH[i].tab); free((void *)m->tabS[i].tab); }
  free(m->tabH); free(m->tabS);
  free((void *)m->nShell); free((void *)m->angShell); free((void *)m->nOrbSpecies);
  free(m);
}

/* SlakoCont_getCutoff, slakocont.F90:139-150: max over all tables */
double orc_model_cutoff(const orc_model *m)
{
  double c = 0.0; int i;
  for (i = 0; i < m->nSpecies * m->nSpecies; ++i) {
    double a = orc_sk_cutoff(m->tabH[i].nGrid, m->tabH[i].dist, m->tabH[i].method);
    double b = orc_sk_cutoff(m->tabS[i].nGrid, m->tabS[i].dist, m->tabS[i].method);
    if (m->tabH[i].tab && a > c) c = a;
    if (m->tabS[i].tab && b > c) c = b;
  }
  return c;
}

static void getSKIntegrals(const orc_model *m, int which, double *sk, double dist, int sp1, int sp2)
{
  const orc_table *t = (which == 0 ? m->tabH : m->tabS) + (sp1 - 1) + m->nSpecies * (sp2 - 1);
  if (t->method == 1) interOld(t, sk, dist); else interNew(t, sk, dist);
}

/* ------------------------------------------------------------------------------------------
 * dftb/sk.F90 : rotation of the diatomic integrals into the molecular frame.
 *
 * The reference (sk.F90:144-945) hard-codes Maple-generated direction-cosine polynomials for
 * every shell pair ss..ff.  Those closed forms are the matrix elements
 *     E(m2, m1) = sum_sigma  D^{l2}(m2,sigma) * D^{l1}(m1,sigma) * V_{|sigma|}
 * where D^l is the (2l+1)x(2l+1) rotation matrix of the Racah-normalised tesseral harmonics
 * (orbital order type/orbitals.F90:59-64: p = y,z,x; d = xy,yz,z2,xz,x2-y2; f = y(3x2-y2),
 * xyz, yz2, z3, xz2, z(x2-y2), x(x2-3y2)) for any rotation taking the z axis onto the bond
 * direction (ll,mm,nn).  The oracle evaluates exactly that definition: D^l by an exact
 * quadrature projection (26-point octahedral rule, exact to degree 7 >= 2*l).  Equality with
 * the reference's polynomials is what tests/test_oracle_sk.py checks against the machine
 * translation of sk.F90 (oracle/_ref) and the golden vectors made from it.
 * ---------------------------------------------------------------------------------------- */

/* Racah-normalised real solid harmonics in the reference's orbital order (read off the
 * sigma-only columns sp/sd/sf, sk.F90:160-240). */
static void tesseral(int l, double x, double y, double z, double *s)
{
  const double r2 = x * x + y * y + z * z;
  switch (l) {
  case 0: s[0] = 1.0; break;
  case 1: s[0] = y; s[1] = z; s[2] = x; break;
  case 2: {
    const double s3 = sqrt(3.0);
    s[0] = s3 * x * y; s[1] = s3 * y * z; s[2] = 1.5 * z * z - 0.5 * r2; s[3] = s3 * x * z;
    s[4] = 0.5 * s3 * (x * x - y * y);
    break;
  }
  default: {
    const double c3 = sqrt(5.0 / 8.0), c2 = sqrt(15.0), c1 = sqrt(3.0 / 8.0);
    s[0] = c3 * y * (3.0 * x * x - y * y);
    s[1] = c2 * x * y * z;
    s[2] = c1 * y * (5.0 * z * z - r2);
    s[3] = 0.5 * z * (5.0 * z * z - 3.0 * r2);
    s[4] = c1 * x * (5.0 * z * z - r2);
    s[5] = 0.5 * c2 * z * (x * x - y * y);
    s[6] = c3 * x * (x * x - 3.0 * y * y);
  }
  }
}

static double quadP[26][3], quadW[26];
static double quadS[ORC_MAXL + 1][26][7];
static int quadReady = 0;

static void quadInit(void)
{
  int n = 0, i, j, k, l;
  const double a = 1.0 / sqrt(2.0), b = 1.0 / sqrt(3.0);
  for (i = 0; i < 3; ++i)
    for (j = -1; j <= 1; j += 2) {
      quadP[n][0] = quadP[n][1] = quadP[n][2] = 0.0; quadP[n][i] = j; quadW[n++] = 1.0 / 21.0;
    }
  for (i = 0; i < 3; ++i)
    for (j = -1; j <= 1; j += 2)
      for (k = -1; k <= 1; k += 2) {
        quadP[n][i] = 0.0; quadP[n][(i + 1) % 3] = a * j; quadP[n][(i + 2) % 3] = a * k;
        quadW[n++] = 4.0 / 105.0;
      }
  for (i = -1; i <= 1; i += 2)
    for (j = -1; j <= 1; j += 2)
      for (k = -1; k <= 1; k += 2) {
        quadP[n][0] = b * i; quadP[n][1] = b * j; quadP[n][2] = b * k; quadW[n++] = 27.0 / 840.0;
      }
  for (l = 0; l <= ORC_MAXL; ++l)
    for (i = 0; i < 26; ++i) tesseral(l, quadP[i][0], quadP[i][1], quadP[i][2], quadS[l][i]);
  quadReady = 1;
}

void orc_init(void)
{
  if (!quadReady) quadInit();
}

/* D^l(m, sigma) = (2l+1) * <S_lm(p) | S_lsigma(R^T p)>_sphere for R = [e1 e2 d]. */
static void rotMatrix(int l, const double R[3][3], double D[7][7])
{
  int q, a, b, n = 2 * l + 1;
  for (a = 0; a < n; ++a) for (b = 0; b < n; ++b) D[a][b] = 0.0;
  for (q = 0; q < 26; ++q) {
    double s[7], px = quadP[q][0], py = quadP[q][1], pz = quadP[q][2];
    /* R^T p : coordinates of p in the bond frame */
    double u = R[0][0] * px + R[1][0] * py + R[2][0] * pz;
    double v = R[0][1] * px + R[1][1] * py + R[2][1] * pz;
    double w = R[0][2] * px + R[1][2] * py + R[2][2] * pz;
    tesseral(l, u, v, w, s);
    for (a = 0; a < n; ++a) {
      double f = quadW[q] * (double)n * quadS[l][q][a];
      for (b = 0; b < n; ++b) D[a][b] += f * s[b];
    }
  }
}

/* rotateH0, sk.F90:35-140.  hh is (ldh, *) column-major, rows = orbitals of species 2. */
static void rotateH0(double *hh, int ldh, const double *skIntegs, double ll, double mm, double nn,
                     int iSp1, int iSp2, const orc_model *orb)
{
  double R[3][3], D[ORC_MAXL + 1][7][7];
  int have[ORC_MAXL + 1] = {0, 0, 0, 0};
  int iSh1, iSh2, ind = 0, iCol = 0, iRow, a, b, s;
  int n1 = orb->nOrbSpecies[iSp1 - 1], n2 = orb->nOrbSpecies[iSp2 - 1];
  /* orthonormal frame with third column = bond direction */
  {
    double sg = nn >= 0.0 ? 1.0 : -1.0;
    double aa = -1.0 / (sg + nn), bb = ll * mm * aa;
    R[0][0] = 1.0 + sg * ll * ll * aa; R[1][0] = sg * bb; R[2][0] = -sg * ll;
    R[0][1] = bb; R[1][1] = sg + mm * mm * aa; R[2][1] = -mm;
    R[0][2] = ll; R[1][2] = mm; R[2][2] = nn;
  }
  for (a = 0; a < n1; ++a) for (b = 0; b < n2; ++b) hh[b + ldh * a] = 0.0;
  for (iSh1 = 0; iSh1 < orb->nShell[iSp1 - 1]; ++iSh1) {
    int ang1 = orb->angShell[iSh1 + orb->mShell * (iSp1 - 1)], nOrb1 = 2 * ang1 + 1;
    if (!have[ang1]) { rotMatrix(ang1, R, D[ang1]); have[ang1] = 1; }
    iRow = 0;
    for (iSh2 = 0; iSh2 < orb->nShell[iSp2 - 1]; ++iSh2) {
      int ang2 = orb->angShell[iSh2 + orb->mShell * (iSp2 - 1)], nOrb2 = 2 * ang2 + 1;
      int lmin = ang1 < ang2 ? ang1 : ang2;
      /* sk.F90:128-133: for ang1 > ang2 the reference evaluates the (ang2,ang1) routine and
       * takes (-1)^(ang1+ang2) * transpose; in the D-matrix form that is one sign factor. */
      double sgn = (ang1 > ang2 && ((ang1 + ang2) & 1)) ? -1.0 : 1.0;
      if (!have[ang2]) { rotMatrix(ang2, R, D[ang2]); have[ang2] = 1; }
      for (a = 0; a < nOrb1; ++a)
        for (b = 0; b < nOrb2; ++b) {
          double e = 0.0;
          for (s = -lmin; s <= lmin; ++s)
            e += D[ang2][b][ang2 + s] * D[ang1][a][ang1 + s] * skIntegs[ind + (s < 0 ? -s : s)];
          hh[(iRow + b) + ldh * (iCol + a)] = sgn * e;
        }
      ind += lmin + 1;
      iRow += nOrb2;
    }
    iCol += nOrb1;
  }
}

/* exposed for the rotation tests: one species pair described by shell lists */
void orc_rotate_h0(const orc_model *m, double *hh, const double *sk, double ll, double mm,
                   double nn, int iSp1, int iSp2)
{
  orc_init();
  rotateH0(hh, m->mOrb, sk, ll, mm, nn, iSp1, iSp2, m);
}

/* ------------------------------------------------------------------------------------------
 * dftb/nonscc.F90 : buildH0 / buildS
 * ---------------------------------------------------------------------------------------- */

/* buildDiatomicBlocks, nonscc.F90:364-424.  iNeighbour/iPair are (0:mNeigh, nAtom). */
static void buildDiatomicBlocks(int firstAtom, int lastAtom, const orc_model *m, int which,
                                const double *coords, const int *nNeighbourSK,
                                const int *iNeighbours, int mNeigh, const int *species,
                                const int *iPair, double *out)
{
  int iAt1;
  const int ld = mNeigh + 1;
#pragma omp parallel for schedule(runtime)
  for (iAt1 = firstAtom; iAt1 <= lastAtom; ++iAt1) {
    double tmp[ORC_MAXORB * ORC_MAXORB], interSK[ORC_MAXINT], vect[3], dist;
    int iSp1 = species[iAt1 - 1], nOrb1 = m->nOrbSpecies[iSp1 - 1], iNeigh1, a, b;
    for (iNeigh1 = 1; iNeigh1 <= nNeighbourSK[iAt1 - 1]; ++iNeigh1) {
      int iAt2 = iNeighbours[iNeigh1 + ld * (iAt1 - 1)];
      int iSp2 = species[iAt2 - 1], nOrb2 = m->nOrbSpecies[iSp2 - 1];
      int ind = iPair[iNeigh1 + ld * (iAt1 - 1)];
      vect[0] = coords[0 + 3 * (iAt2 - 1)] - coords[0 + 3 * (iAt1 - 1)];
      vect[1] = coords[1 + 3 * (iAt2 - 1)] - coords[1 + 3 * (iAt1 - 1)];
      vect[2] = coords[2 + 3 * (iAt2 - 1)] - coords[2 + 3 * (iAt1 - 1)];
      dist = sqrt(vect[0] * vect[0] + vect[1] * vect[1] + vect[2] * vect[2]);
      vect[0] /= dist; vect[1] /= dist; vect[2] /= dist;
      getSKIntegrals(m, which, interSK, dist, iSp1, iSp2);
      rotateH0(tmp, m->mOrb, interSK, vect[0], vect[1], vect[2], iSp1, iSp2, m);
      for (a = 0; a < nOrb1; ++a)
        for (b = 0; b < nOrb2; ++b) out[ind + b + nOrb2 * a] = tmp[b + m->mOrb * a];
    }
  }
}

/* getChunkRanges, common/schedule.F90:230-263 */
void orc_chunk_ranges(int groupSize, int myRank, int globalFirst, int globalLast, int *localFirst,
                      int *localLast)
{
  int rangeLength = globalLast - globalFirst + 1;
  int nLocal = rangeLength / groupSize;
  int remainder = rangeLength % groupSize;
  if (myRank < remainder) {
    nLocal += 1;
    *localFirst = globalFirst + myRank * nLocal;
  } else {
    *localFirst = globalFirst + remainder * (nLocal + 1) + (myRank - remainder) * nLocal;
  }
  *localLast = *localFirst + nLocal - 1 < globalLast ? *localFirst + nLocal - 1 : globalLast;
}

/* getIndicesWithWorkload, common/schedule.F90:267-307; returns the count, indices 1-based */
int orc_indices_with_workload(int groupSize, int myRank, int globalFirst, int globalLast,
                              const int *workload, int *indices)
{
  int *rankWorkload = (int *)calloc(groupSize, sizeof(int));
  int numIndices = 0, ii, r, rank;
  for (ii = globalFirst; ii <= globalLast; ++ii) {
    rank = 0;
    for (r = 1; r < groupSize; ++r) if (rankWorkload[r] < rankWorkload[rank]) rank = r;
    rankWorkload[rank] += workload[ii - 1] > 1 ? workload[ii - 1] : 1;
    if (rank == myRank) indices[numIndices++] = ii;
  }
  free(rankWorkload);
  return numIndices;
}

/* buildH0 (selfegy != NULL, nonscc.F90:66-124) and buildS (selfegy == NULL, :128-181).
 * selfegy is (mShell, nSpecies).  rank/nRanks restate distributeRangeInChunks
 * (schedule.F90:51-76); the caller sums the per-rank arrays (assembleChunks :170-182). */
void orc_build_sparse(const orc_model *m, const double *selfegy, int nAtom, const double *coords,
                      const int *nNeighbourSK, const int *iNeighbours, int mNeigh,
                      const int *species, const int *iPair, int sparseSize, double *out,
                      int rank, int nRanks)
{
  int iAtFirst, iAtLast, iAt1, i;
  const int ld = mNeigh + 1;
  orc_init();
  for (i = 0; i < sparseSize; ++i) out[i] = 0.0;
  orc_chunk_ranges(nRanks, rank, 1, nAtom, &iAtFirst, &iAtLast);
  for (iAt1 = iAtFirst; iAt1 <= iAtLast; ++iAt1) {
    int iSp1 = species[iAt1 - 1], n = m->nOrbSpecies[iSp1 - 1];
    int ind = iPair[0 + ld * (iAt1 - 1)], iSh, k;
    for (iSh = 0; iSh < m->nShell[iSp1 - 1]; ++iSh) {
      int l = m->angShell[iSh + m->mShell * (iSp1 - 1)];
      for (k = 0; k < 2 * l + 1; ++k) {
        out[ind] = selfegy ? selfegy[iSh + m->mShell * (iSp1 - 1)] : 1.0;
        ind += n + 1;
      }
    }
  }
  buildDiatomicBlocks(iAtFirst, iAtLast, m, selfegy ? 0 : 1, coords, nNeighbourSK, iNeighbours,
                      mNeigh, species, iPair, out);
}

/* getFirstDerivFiniteDiff, nonscc.F90:428-484.  deriv is (mOrb, mOrb, 3). */
static void firstDerivFiniteDiff(double *deriv, const orc_model *m, int which, const double *coords,
                                 const int *species, int atomI, int atomJ, double deltaXDiff)
{
  const int mo = m->mOrb, sz = mo * mo;
  double tmp[2][3][ORC_MAXORB * ORC_MAXORB], interSk[ORC_MAXINT], vect[3], dist;
  int ii, jj, k, sp1 = species[atomI - 1], sp2 = species[atomJ - 1];
  for (jj = 0; jj < 3; ++jj)
    for (ii = 1; ii <= 2; ++ii) {
      for (k = 0; k < 3; ++k) vect[k] = coords[k + 3 * (atomJ - 1)] - coords[k + 3 * (atomI - 1)];
      vect[jj] = vect[jj] - (double)(2 * ii - 3) * deltaXDiff;
      dist = sqrt(vect[0] * vect[0] + vect[1] * vect[1] + vect[2] * vect[2]);
      vect[0] /= dist; vect[1] /= dist; vect[2] /= dist;
      for (k = 0; k < sz; ++k) tmp[ii - 1][jj][k] = 0.0;
      getSKIntegrals(m, which, interSk, dist, sp1, sp2);
      rotateH0(tmp[ii - 1][jj], mo, interSk, vect[0], vect[1], vect[2], sp1, sp2, m);
    }
  for (ii = 0; ii < 3; ++ii)
    for (k = 0; k < sz; ++k)
      deriv[k + sz * ii] = (tmp[1][ii][k] - tmp[0][ii][k]) / (2.0 * deltaXDiff);
}

/* getFirstDerivRichardson, nonscc.F90:488-600 */
static void firstDerivRichardson(double *deriv, const orc_model *m, int which, const double *coords,
                                 const int *species, int atomI, int atomJ)
{
  enum { maxrows = 20, maxcolumns = 3 };
  const double tol = 1.0e-11;
  const int mo = m->mOrb, sz = mo * mo;
  const double hmin = pow(2.220446049250313e-16, 0.25);
  static __thread double dd[maxrows + 1][maxcolumns + 1][ORC_MAXORB * ORC_MAXORB];
  double tmp[2][ORC_MAXORB * ORC_MAXORB], interSk[ORC_MAXINT], vect[3], dist, hh;
  char conv[ORC_MAXORB * ORC_MAXORB];
  int iCart, kk, k, ii, mm = 0, r, c, sp1 = species[atomI - 1], sp2 = species[atomJ - 1];
  for (iCart = 0; iCart < 3; ++iCart) {
    hh = 0.125;
    memset(dd, 0, sizeof(dd));
    for (kk = 1; kk <= 2; ++kk) {
      for (k = 0; k < 3; ++k) vect[k] = coords[k + 3 * (atomJ - 1)] - coords[k + 3 * (atomI - 1)];
      vect[iCart] = vect[iCart] - (double)(2 * kk - 3) * hh;
      dist = sqrt(vect[0] * vect[0] + vect[1] * vect[1] + vect[2] * vect[2]);
      vect[0] /= dist; vect[1] /= dist; vect[2] /= dist;
      for (k = 0; k < sz; ++k) tmp[kk - 1][k] = 0.0;
      getSKIntegrals(m, which, interSk, dist, sp1, sp2);
      rotateH0(tmp[kk - 1], mo, interSk, vect[0], vect[1], vect[2], sp1, sp2, m);
    }
    ii = 0; mm = 0;
    for (k = 0; k < sz; ++k) { dd[0][0][k] = (tmp[1][k] - tmp[0][k]) / (2.0 * hh); conv[k] = 0; }
    for (;;) {
      int all = 1;
      for (k = 0; k < sz; ++k) all = all && conv[k];
      if (!(ii < maxrows && !all)) break;
      hh = hh / 2.0;
      if (hh <= hmin) break;
      ii += 1;
      for (kk = 1; kk <= 2; ++kk) {
        for (k = 0; k < 3; ++k) vect[k] = coords[k + 3 * (atomJ - 1)] - coords[k + 3 * (atomI - 1)];
        vect[iCart] = vect[iCart] - (double)(2 * kk - 3) * hh;
        dist = sqrt(vect[0] * vect[0] + vect[1] * vect[1] + vect[2] * vect[2]);
        vect[0] /= dist; vect[1] /= dist; vect[2] /= dist;
        for (k = 0; k < sz; ++k) tmp[kk - 1][k] = 0.0;
        getSKIntegrals(m, which, interSk, dist, sp1, sp2);
        rotateH0(tmp[kk - 1], mo, interSk, vect[0], vect[1], vect[2], sp1, sp2, m);
      }
      for (k = 0; k < sz; ++k)
        if (!conv[k]) dd[ii][0][k] = (tmp[1][k] - tmp[0][k]) / (2.0 * hh);
      mm = ii < maxcolumns ? ii : maxcolumns;
      for (kk = 1; kk <= mm; ++kk) {
        int nk = 2 * kk;
        double p = (double)(1 << nk);
        for (k = 0; k < sz; ++k)
          if (!conv[k]) dd[ii][kk][k] = (p * dd[ii][kk - 1][k] - dd[ii - 1][kk - 1][k]) / (p - 1.0);
      }
      for (k = 0; k < sz; ++k) {
        double diff;
        if (conv[k]) continue;
        diff = fabs(dd[ii][mm - 1][k] - dd[ii][mm][k]);
        if (diff <= tol * dd[ii][mm][k]) {
          double v = dd[ii][mm][k];
          for (r = 0; r <= maxrows; ++r) for (c = 0; c <= maxcolumns; ++c) dd[r][c][k] = v;
          conv[k] = 1;
        }
      }
    }
    for (k = 0; k < sz; ++k) deriv[k + sz * iCart] = dd[ii][mm][k];
  }
}

/* TNonSccDiff%getFirstDerivBlock, nonscc.F90:210-244; diffType 1 = finiteDiff, 2 = richardson */
void orc_first_deriv_block(const orc_model *m, int which, double *deriv, const double *coords,
                           const int *species, int atomI, int atomJ, int diffType, double delta)
{
  orc_init();
  if (diffType == 2) firstDerivRichardson(deriv, m, which, coords, species, atomI, atomJ);
  else firstDerivFiniteDiff(deriv, m, which, coords, species, atomI, atomJ, delta);
}

/* ------------------------------------------------------------------------------------------
 * dftb/periodic.F90 : neighbour list;  dftb/sparse2dense.F90 : sparse descriptor
 * ---------------------------------------------------------------------------------------- */

/* index_heap_sort with tolerance, math/sorting.F90:142-199 (1-based indx) */
static void index_heap_sort(int n, int *indx, const double *array, double tol)
{
  int ii, ir, ij, il, indxTmp;
  double arrayTmp;
  for (ii = 1; ii <= n; ++ii) indx[ii - 1] = ii;
  if (n <= 1) return;
  il = n / 2 + 1;
  ir = n;
  for (;;) {
    if (il > 1) {
      il -= 1;
      indxTmp = indx[il - 1];
      arrayTmp = array[indxTmp - 1];
    } else {
      indxTmp = indx[ir - 1];
      arrayTmp = array[indxTmp - 1];
      indx[ir - 1] = indx[0];
      ir -= 1;
      if (ir < 1) { indx[0] = indxTmp; return; }
    }
    ii = il;
    ij = 2 * il;
    while (ij <= ir) {
      if (ij < ir) {
        if (array[indx[ij - 1] - 1] < array[indx[ij] - 1] - tol) ij += 1;
      }
      if (arrayTmp < array[indx[ij - 1] - 1] - tol) {
        indx[ii - 1] = indx[ij - 1];
        ii = ij;
        ij = 2 * ij;
      } else {
        ij = ir + 1;
      }
    }
    indx[ii - 1] = indxTmp;
  }
}

/* foldCoordsToCell, pbc3d branch without contacts (dftb/boundarycond.F90:369-380), with invert33 =
 * invertInOut + determinant33 (math/simplealgebra.F90:45-66 and the cofactor form that follows).
 * coord(3,nAtom) in/out, latVec(3,3) column-major (column i = i-th lattice vector). */
void orc_fold_coords_to_cell(int nAtom, double *coord, const double *latVec)
{
#define M(i, j) latVec[((i) - 1) + 3 * ((j) - 1)]
#define INV(i, j) inv[((i) - 1) + 3 * ((j) - 1)]
  double inv[9], vecLen[3], det, tmp;
  int iAt, i, j;
  tmp = M(1, 1) * (M(2, 2) * M(3, 3) - M(3, 2) * M(2, 3));
  tmp = tmp - M(1, 2) * (M(2, 1) * M(3, 3) - M(3, 1) * M(2, 3));
  tmp = tmp + M(1, 3) * (M(2, 1) * M(3, 2) - M(3, 1) * M(2, 2));
  det = tmp;
  INV(1, 1) = -M(2, 3) * M(3, 2) + M(2, 2) * M(3, 3);
  INV(2, 1) = M(2, 3) * M(3, 1) - M(2, 1) * M(3, 3);
  INV(3, 1) = -M(2, 2) * M(3, 1) + M(2, 1) * M(3, 2);
  INV(1, 2) = M(1, 3) * M(3, 2) - M(1, 2) * M(3, 3);
  INV(2, 2) = -M(1, 3) * M(3, 1) + M(1, 1) * M(3, 3);
  INV(3, 2) = M(1, 2) * M(3, 1) - M(1, 1) * M(3, 2);
  INV(1, 3) = -M(1, 3) * M(2, 2) + M(1, 2) * M(2, 3);
  INV(2, 3) = M(1, 3) * M(2, 1) - M(1, 1) * M(2, 3);
  INV(3, 3) = -M(1, 2) * M(2, 1) + M(1, 1) * M(2, 2);
  for (i = 0; i < 9; ++i) inv[i] = inv[i] / det;
  for (j = 1; j <= 3; ++j) vecLen[j - 1] = sqrt(M(1, j) * M(1, j) + M(2, j) * M(2, j) + M(3, j) * M(3, j));
  for (iAt = 0; iAt < nAtom; ++iAt) {
    double *x = coord + 3 * iAt, frac[3], out[3];
    for (i = 1; i <= 3; ++i) frac[i - 1] = INV(i, 1) * x[0] + INV(i, 2) * x[1] + INV(i, 3) * x[2];
    for (i = 0; i < 3; ++i) {
      frac[i] = frac[i] - floor(frac[i]);
      if (fabs(vecLen[i] * (1.0 - frac[i])) < DBL_EPSILON) frac[i] = 0.0;
    }
    for (i = 1; i <= 3; ++i) out[i - 1] = M(i, 1) * frac[0] + M(i, 2) * frac[1] + M(i, 3) * frac[2];
    x[0] = out[0]; x[1] = out[1]; x[2] = out[2];
  }
#undef M
#undef INV
}

/* getCellTranslations -> getLatticePoints -> TLatPointIter (periodic.F90:227-263,272-336;
 * type/latpointiter.F90:62-126,130-184,214-274) with posExtension = negExtension = 1 and all
 * points of the parallelepiped kept (tAll).  latVec(3,3), recVec2p(3,3) column-major, the
 * latter holding the reciprocal vectors in units of 2 pi.  Returns nCellVec; cellVec may be
 * NULL to query the count. */
int orc_cell_translations(const double *latVec, const double *recVec2p, double cutoff,
                          double *cellVec, double *rCellVec)
{
  int rng[3], i, a, b, c, n = 0, k;
  for (i = 0; i < 3; ++i) {
    const double *v = recVec2p + 3 * i;
    rng[i] = (int)floor(cutoff * sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2])) + 1;
  }
  if (cellVec) {
    for (k = 0; k < 3; ++k) { cellVec[k] = 0.0; rCellVec[k] = 0.0; }
  }
  n = 1;
  for (a = -rng[0]; a <= rng[0]; ++a)
    for (b = -rng[1]; b <= rng[1]; ++b)
      for (c = -rng[2]; c <= rng[2]; ++c) {
        if (a == 0 && b == 0 && c == 0) continue;
        if (cellVec) {
          cellVec[3 * n + 0] = a; cellVec[3 * n + 1] = b; cellVec[3 * n + 2] = c;
          for (k = 0; k < 3; ++k)
            rCellVec[3 * n + k] = latVec[k + 0] * a + latVec[k + 3] * b + latVec[k + 6] * c;
        }
        ++n;
      }
  return n;
}

/* updateNeighbourList (isSymm = 0: the non-symmetric half list; 1: the symmetric one), periodic.F90:446-694, followed by
 * getNrOfNeighboursForAll (:996-1054) for the SK cutoff.  Outputs sized by the caller:
 *   coord(3,maxAllAtom), img2CentCell/iCellVec(maxAllAtom), iNeighbour(0:maxNeigh, nAtom),
 *   neighDist2(0:maxNeigh, nAtom).
 * Returns 0 ok; 1 = maxAllAtom too small; 2 = maxNeigh too small; -1 = atoms too close
 * (pairError).  nAllAtom / mNeigh report the needed sizes either way. */
int orc_update_neighbour_list(int nAtom, const double *coord0, double cutoff, int nCellVec,
                              const double *rCellVec, int maxAllAtom, int maxNeigh, double *coord,
                              int *img2CentCell, int *iCellVec, int *nNeighbour, int *iNeighbour,
                              double *neighDist2, int *nAllAtomOut, int *mNeighOut, int *pairError, int isSymm)
{
  const double cutoff2 = cutoff * cutoff;
  const int ld = maxNeigh + 1;
  int ii, iAtom1, iAtom2, oldIAtom1, nAllAtom = 0, k, overflowA = 0, overflowN = 0, mN = 0;
  int *indx, *itmp;
  double *dtmp;
  pairError[0] = pairError[1] = 0;
  for (k = 0; k < nAtom; ++k) nNeighbour[k] = 0;
  for (k = 0; k < ld * nAtom; ++k) { iNeighbour[k] = 0; neighDist2[k] = 0.0; }
  for (ii = 1; ii <= nCellVec; ++ii) {
    const double *rCell = rCellVec + 3 * (ii - 1);
    oldIAtom1 = 0;
    for (iAtom1 = 1; iAtom1 <= nAtom; ++iAtom1) {
      double rr[3];
      for (k = 0; k < 3; ++k) rr[k] = coord0[k + 3 * (iAtom1 - 1)] + rCell[k];
      /* iAtom2End = nAtom for the symmetric list, iAtom1 otherwise (periodic.F90:580-584) */
      for (iAtom2 = 1; iAtom2 <= (isSymm ? nAtom : iAtom1); ++iAtom2) {
        double d0 = coord0[0 + 3 * (iAtom2 - 1)] - rr[0];
        double d1 = coord0[1 + 3 * (iAtom2 - 1)] - rr[1];
        double d2 = coord0[2 + 3 * (iAtom2 - 1)] - rr[2];
        double dist2 = d0 * d0 + d1 * d1 + d2 * d2;
        if (dist2 > cutoff2) continue;
        if (iAtom1 != oldIAtom1) {
          nAllAtom += 1;
          if (nAllAtom > maxAllAtom) overflowA = 1;
          else {
            for (k = 0; k < 3; ++k) coord[k + 3 * (nAllAtom - 1)] = rr[k];
            img2CentCell[nAllAtom - 1] = iAtom1;
            iCellVec[nAllAtom - 1] = ii;
          }
          oldIAtom1 = iAtom1;
        }
        if (dist2 < minNeighDist2) {
          if (ii == 1 && iAtom1 == iAtom2) continue;
          pairError[0] = iAtom1; pairError[1] = iAtom2;
          return -1;
        }
        nNeighbour[iAtom2 - 1] += 1;
        if (nNeighbour[iAtom2 - 1] > maxNeigh) overflowN = 1;
        else {
          iNeighbour[nNeighbour[iAtom2 - 1] + ld * (iAtom2 - 1)] = nAllAtom;
          neighDist2[nNeighbour[iAtom2 - 1] + ld * (iAtom2 - 1)] = dist2;
        }
      }
    }
  }
  for (k = 0; k < nAtom; ++k) if (nNeighbour[k] > mN) mN = nNeighbour[k];
  *nAllAtomOut = nAllAtom; *mNeighOut = mN;
  if (overflowA) return 1;
  if (overflowN) return 2;
  /* sort by distance, periodic.F90:675-688; row 0 = the atom itself (:fillNeighbourArrays) */
  indx = (int *)malloc(sizeof(int) * (mN + 1));
  itmp = (int *)malloc(sizeof(int) * (mN + 1));
  dtmp = (double *)malloc(sizeof(double) * (mN + 1));
  for (iAtom1 = 1; iAtom1 <= nAtom; ++iAtom1) {
    int nn1 = nNeighbour[iAtom1 - 1];
    int *in = iNeighbour + ld * (iAtom1 - 1) + 1;
    double *dn = neighDist2 + ld * (iAtom1 - 1) + 1;
    index_heap_sort(nn1, indx, dn, tolSameDist2);
    for (k = 0; k < nn1; ++k) { itmp[k] = in[indx[k] - 1]; dtmp[k] = dn[indx[k] - 1]; }
    for (k = 0; k < nn1; ++k) { in[k] = itmp[k]; dn[k] = dtmp[k]; }
    iNeighbour[0 + ld * (iAtom1 - 1)] = iAtom1;
    neighDist2[0 + ld * (iAtom1 - 1)] = 0.0;
  }
  free(indx); free(itmp); free(dtmp);
  return 0;
}

/* search_asc_real_geq, math/binarysearch.F90:63-94: last jj with xVal - xx(jj) >= -tol on an ascending
 * array xx(1:n) (0 if none).  xx points at xx(1). */
int orc_search_asc_real_geq(int n, const double *xx, double xVal, double tol)
{
  int jlower = 0, jupper = n, jcurr;
  while (jlower < jupper) {
    jcurr = jlower + (jupper - jlower + 1) / 2;
    if (xVal - xx[jcurr - 1] >= -tol) jlower = jcurr; else jupper = jcurr - 1;
  }
  return jlower;
}

/* index_heap_sort (math/sorting.F90:142-199) for the tests: indx(n) 1-based on return */
void orc_index_heap_sort(int n, int *indx, const double *array, double tol)
{
  index_heap_sort(n, indx, array, tol);
}

/* getNrOfNeighbours, periodic.F90:1026-1054: search_asc_real_geq(nNeighbour, neighDist2(1:nNeighbour(iAtom),
 * iAtom), cutoff**2, tolSameDist2).  (The bisection is kept literally: the list is only ascending
 * within tolSameDist2.) */
void orc_nr_of_neighbours(int nAtom, const int *nNeighbour, const double *neighDist2, int maxNeigh,
                          double cutoff, int *nNeighbourSK)
{
  const int ld = maxNeigh + 1;
  const double c2 = cutoff * cutoff;
  int iAtom;
  for (iAtom = 0; iAtom < nAtom; ++iAtom)
    nNeighbourSK[iAtom] = orc_search_asc_real_geq(nNeighbour[iAtom], neighDist2 + ld * iAtom + 1, c2, tolSameDist2);
}

/* getSparseDescriptor, sparse2dense.F90:3617-3662; returns sparseSize.  nOrbAtom(nAtom). */
int orc_sparse_descriptor(int nAtom, const int *iNeighbour, int mNeigh, const int *nNeighbour,
                          const int *img2CentCell, const int *nOrbAtom, int *iPair)
{
  const int ld = mNeigh + 1;
  int ind = 0, iAt1, iNeigh1;
  for (iAt1 = 0; iAt1 < ld * nAtom; ++iAt1) iPair[iAt1] = 0;
  for (iAt1 = 1; iAt1 <= nAtom; ++iAt1) {
    int nOrb1 = nOrbAtom[iAt1 - 1];
    for (iNeigh1 = 0; iNeigh1 <= nNeighbour[iAt1 - 1]; ++iNeigh1) {
      int nOrb2 = nOrbAtom[img2CentCell[iNeighbour[iNeigh1 + ld * (iAt1 - 1)] - 1] - 1];
      iPair[iNeigh1 + ld * (iAt1 - 1)] = ind;
      ind += nOrb1 * nOrb2;
    }
  }
  return ind;
}

/* ------------------------------------------------------------------------------------------
 * dftb/sparse2dense.F90 : unpackHS / packHS (serial, as in the reference)
 * ---------------------------------------------------------------------------------------- */

/* unpackHS_real, sparse2dense.F90:266-321.  square(nOrb,nOrb), iAtomStart(nAtom+1) 1-based. */
void orc_unpack_hs_real(double *square, int nOrb, const double *orig, int nAtom,
                        const int *iNeighbour, int mNeigh, const int *nNeighbourSK,
                        const int *iAtomStart, const int *iSparseStart, const int *img2CentCell)
{
  const int ld = mNeigh + 1;
  int iAtom1, iNeigh, a, b;
  memset(square, 0, sizeof(double) * (size_t)nOrb * nOrb);
  for (iAtom1 = 1; iAtom1 <= nAtom; ++iAtom1) {
    int ii = iAtomStart[iAtom1 - 1], nOrb1 = iAtomStart[iAtom1] - ii;
    for (iNeigh = 0; iNeigh <= nNeighbourSK[iAtom1 - 1]; ++iNeigh) {
      int iOrig = iSparseStart[iNeigh + ld * (iAtom1 - 1)];
      int iAtom2f = img2CentCell[iNeighbour[iNeigh + ld * (iAtom1 - 1)] - 1];
      int jj = iAtomStart[iAtom2f - 1], nOrb2 = iAtomStart[iAtom2f] - jj;
      for (a = 0; a < nOrb1; ++a)
        for (b = 0; b < nOrb2; ++b)
          square[(jj - 1 + b) + (size_t)nOrb * (ii - 1 + a)] += orig[iOrig + b + nOrb2 * a];
    }
  }
}

/* unpackHS_cmplx_kpts, sparse2dense.F90:185-260.  square is complex (re,im interleaved). */
void orc_unpack_hs_cplx(double *square, int nOrb, const double *orig, const double *kPoint,
                        int nAtom, const int *iNeighbour, int mNeigh, const int *nNeighbourSK,
                        const int *iCellVec, const double *cellVec, const int *iAtomStart,
                        const int *iSparseStart, const int *img2CentCell)
{
  const int ld = mNeigh + 1;
  const double pi = 3.14159265358979323846264338327950288;
  double k2p[3] = {2.0 * pi * kPoint[0], 2.0 * pi * kPoint[1], 2.0 * pi * kPoint[2]};
  double pr = 1.0, pim = 0.0;
  int iOldVec = 0, iAtom1, iNeigh, a, b;
  memset(square, 0, sizeof(double) * 2 * (size_t)nOrb * nOrb);
  for (iAtom1 = 1; iAtom1 <= nAtom; ++iAtom1) {
    int ii = iAtomStart[iAtom1 - 1], nOrb1 = iAtomStart[iAtom1] - ii;
    for (iNeigh = 0; iNeigh <= nNeighbourSK[iAtom1 - 1]; ++iNeigh) {
      int iOrig = iSparseStart[iNeigh + ld * (iAtom1 - 1)];
      int iAtom2 = iNeighbour[iNeigh + ld * (iAtom1 - 1)];
      int iAtom2f = img2CentCell[iAtom2 - 1];
      int jj = iAtomStart[iAtom2f - 1], nOrb2 = iAtomStart[iAtom2f] - jj;
      int iVec = iCellVec[iAtom2 - 1];
      if (iVec != iOldVec) {
        const double *cv = cellVec + 3 * (iVec - 1);
        double ph = k2p[0] * cv[0] + k2p[1] * cv[1] + k2p[2] * cv[2];
        pr = cos(ph); pim = sin(ph);
        iOldVec = iVec;
      }
      for (a = 0; a < nOrb1; ++a)
        for (b = 0; b < nOrb2; ++b) {
          size_t p = 2 * ((size_t)(jj - 1 + b) + (size_t)nOrb * (ii - 1 + a));
          double v = orig[iOrig + b + nOrb2 * a];
          square[p] += pr * v;
          square[p + 1] += pim * v;
        }
    }
  }
}

/* packHS_real, sparse2dense.F90:1024-1096 (accumulates into primitive) */
void orc_pack_hs_real(double *primitive, const double *square, int nOrb, int nAtom,
                      const int *iNeighbour, int mNeigh, const int *nNeighbourSK,
                      const int *iAtomStart, const int *iSparseStart, const int *img2CentCell)
{
  const int ld = mNeigh + 1;
  double tmpSqr[ORC_MAXORB * ORC_MAXORB];
  int iAtom1, iNeigh, a, b, kk;
  for (iAtom1 = 1; iAtom1 <= nAtom; ++iAtom1) {
    int ii = iAtomStart[iAtom1 - 1], nOrb1 = iAtomStart[iAtom1] - ii;
    for (iNeigh = 0; iNeigh <= nNeighbourSK[iAtom1 - 1]; ++iNeigh) {
      int iOrig = iSparseStart[iNeigh + ld * (iAtom1 - 1)];
      int iAtom2f = img2CentCell[iNeighbour[iNeigh + ld * (iAtom1 - 1)] - 1];
      int jj = iAtomStart[iAtom2f - 1], nOrb2 = iAtomStart[iAtom2f] - jj;
      for (a = 0; a < nOrb1; ++a)
        for (b = 0; b < nOrb2; ++b)
          tmpSqr[b + ORC_MAXORB * a] = square[(jj - 1 + b) + (size_t)nOrb * (ii - 1 + a)];
      if (iAtom1 == iAtom2f)
        for (kk = 0; kk < nOrb2; ++kk)
          for (a = kk + 1; a < nOrb1; ++a) tmpSqr[kk + ORC_MAXORB * a] = tmpSqr[a + ORC_MAXORB * kk];
      for (a = 0; a < nOrb1; ++a)
        for (b = 0; b < nOrb2; ++b) primitive[iOrig + b + nOrb2 * a] += tmpSqr[b + ORC_MAXORB * a];
    }
  }
}

/* packHS_cmplx_kpts, sparse2dense.F90:820-918 */
void orc_pack_hs_cplx(double *primitive, const double *square, int nOrb, const double *kPoint,
                      double kWeight, int nAtom, const int *iNeighbour, int mNeigh,
                      const int *nNeighbourSK, const int *iCellVec, const double *cellVec,
                      const int *iAtomStart, const int *iSparseStart, const int *img2CentCell)
{
  const int ld = mNeigh + 1;
  const double pi = 3.14159265358979323846264338327950288;
  double k2p[3] = {2.0 * pi * kPoint[0], 2.0 * pi * kPoint[1], 2.0 * pi * kPoint[2]};
  double tr[ORC_MAXORB * ORC_MAXORB], ti[ORC_MAXORB * ORC_MAXORB];
  double pr = 1.0, pim = 0.0;
  int iOldVec = 0, iAtom1, iNeigh, a, b, kk;
  for (iAtom1 = 1; iAtom1 <= nAtom; ++iAtom1) {
    int ii = iAtomStart[iAtom1 - 1], nOrb1 = iAtomStart[iAtom1] - ii;
    for (iNeigh = 0; iNeigh <= nNeighbourSK[iAtom1 - 1]; ++iNeigh) {
      int iOrig = iSparseStart[iNeigh + ld * (iAtom1 - 1)];
      int iAtom2 = iNeighbour[iNeigh + ld * (iAtom1 - 1)];
      int iAtom2f = img2CentCell[iAtom2 - 1];
      int jj = iAtomStart[iAtom2f - 1], nOrb2 = iAtomStart[iAtom2f] - jj;
      int iVec = iCellVec[iAtom2 - 1];
      if (iVec != iOldVec) {
        const double *cv = cellVec + 3 * (iVec - 1);
        double ph = k2p[0] * cv[0] + k2p[1] * cv[1] + k2p[2] * cv[2];
        pr = cos(ph); pim = -sin(ph); /* exp(-i k.R) */
        iOldVec = iVec;
      }
      for (a = 0; a < nOrb1; ++a)
        for (b = 0; b < nOrb2; ++b) {
          size_t p = 2 * ((size_t)(jj - 1 + b) + (size_t)nOrb * (ii - 1 + a));
          tr[b + ORC_MAXORB * a] = square[p];
          ti[b + ORC_MAXORB * a] = square[p + 1];
        }
      if (iAtom1 == iAtom2f)
        for (kk = 0; kk < nOrb2; ++kk)
          for (a = kk + 1; a < nOrb1; ++a) {
            tr[kk + ORC_MAXORB * a] = tr[a + ORC_MAXORB * kk];
            ti[kk + ORC_MAXORB * a] = -ti[a + ORC_MAXORB * kk];
          }
      for (a = 0; a < nOrb1; ++a)
        for (b = 0; b < nOrb2; ++b)
          primitive[iOrig + b + nOrb2 * a] +=
              kWeight * (pr * tr[b + ORC_MAXORB * a] - pim * ti[b + ORC_MAXORB * a]);
    }
  }
}

/* ------------------------------------------------------------------------------------------
 * dftb/populations.F90 : Mulliken
 * ---------------------------------------------------------------------------------------- */

/* addShift_atom, dftb/shift.F90:41-127.  ham(sparseSize, nSpin) column-major, shift(nAtom, nSpin).
 * isInputZero: ham is zeroed first (then the reference distributes the atoms by workload and
 * sums the ranks' arrays: iterIndices restates that; NULL = all atoms). */
void orc_add_shift_atom(double *ham, const double *over, int sparseSize, int nSpin, int nAtom,
                        const int *nOrbAtom, const int *iNeighbour, int mNeigh,
                        const int *nNeighbourSK, const int *img2CentCell, const int *iPair,
                        const double *shift, int isInputZero, const int *iterIndices, int nIter)
{
  const int ld = mNeigh + 1;
  int iSpin, iIter, iNeigh, k;
  if (isInputZero) memset(ham, 0, sizeof(double) * (size_t)sparseSize * nSpin);
  if (!iterIndices) nIter = nAtom;
  for (iSpin = 0; iSpin < nSpin; ++iSpin)
    for (iIter = 0; iIter < nIter; ++iIter) {
      int iAt1 = iterIndices ? iterIndices[iIter] : iIter + 1;
      int nOrb1 = nOrbAtom[iAt1 - 1];
      for (iNeigh = 0; iNeigh <= nNeighbourSK[iAt1 - 1]; ++iNeigh) {
        int iAt2f = img2CentCell[iNeighbour[iNeigh + ld * (iAt1 - 1)] - 1];
        int nOrb2 = nOrbAtom[iAt2f - 1];
        int iOrig = iPair[iNeigh + ld * (iAt1 - 1)];
        double *h = ham + (size_t)sparseSize * iSpin + iOrig;
        for (k = 0; k < nOrb2 * nOrb1; ++k)
          h[k] = h[k] + over[iOrig + k] * 0.5 *
                            (shift[(iAt1 - 1) + nAtom * iSpin] + shift[(iAt2f - 1) + nAtom * iSpin]);
      }
    }
}

/* addShift_block, dftb/shift.F90:132-228.  shift(mOrb, mOrb, nAtom, nSpin). */
void orc_add_shift_block(double *ham, const double *over, int sparseSize, int nSpin, int nAtom,
                         int mOrb, const int *nOrbAtom, const int *iNeighbour, int mNeigh,
                         const int *nNeighbourSK, const int *img2CentCell, const int *iPair,
                         const double *shift, int isInputZero, const int *iterIndices, int nIter)
{
  const int ld = mNeigh + 1;
  int iSpin, iIter, iNeigh, a, b, c;
  if (isInputZero) memset(ham, 0, sizeof(double) * (size_t)sparseSize * nSpin);
  if (!iterIndices) nIter = nAtom;
  for (iSpin = 0; iSpin < nSpin; ++iSpin)
    for (iIter = 0; iIter < nIter; ++iIter) {
      int iAt1 = iterIndices ? iterIndices[iIter] : iIter + 1;
      int nOrb1 = nOrbAtom[iAt1 - 1];
      const double *V1 = shift + (size_t)mOrb * mOrb * ((iAt1 - 1) + (size_t)nAtom * iSpin);
      for (iNeigh = 0; iNeigh <= nNeighbourSK[iAt1 - 1]; ++iNeigh) {
        int iAt2f = img2CentCell[iNeighbour[iNeigh + ld * (iAt1 - 1)] - 1];
        int nOrb2 = nOrbAtom[iAt2f - 1];
        int iOrig = iPair[iNeigh + ld * (iAt1 - 1)];
        const double *V2 = shift + (size_t)mOrb * mOrb * ((iAt2f - 1) + (size_t)nAtom * iSpin);
        const double *S = over + iOrig;
        double *h = ham + (size_t)sparseSize * iSpin + iOrig;
        /* tmpH = 0.5 * (matmul(tmpS, shift(:,:,iAt1)) + matmul(shift(:,:,iAt2f), tmpS)) */
        for (a = 0; a < nOrb1; ++a)
          for (b = 0; b < nOrb2; ++b) {
            double x = 0.0, y = 0.0;
            for (c = 0; c < nOrb1; ++c) x += S[b + nOrb2 * c] * V1[c + mOrb * a];
            for (c = 0; c < nOrb2; ++c) y += V2[b + mOrb * c] * S[c + nOrb2 * a];
            h[b + nOrb2 * a] = h[b + nOrb2 * a] + 0.5 * (x + y);
          }
      }
    }
}

/* mullikenPerOrbital, populations.F90:120-187.  qq(mOrb, nAtom) accumulates.  iterIndices /
 * nIter restate distributeRangeWithWorkload (schedule.F90:137-164); NULL = all atoms. */
void orc_mulliken_per_orbital(double *qq, int mOrb, const double *over, const double *rho,
                              int nAtom, const int *nOrbAtom, const int *iNeighbour, int mNeigh,
                              const int *nNeighbourSK, const int *img2CentCell, const int *iPair,
                              const int *iterIndices, int nIter)
{
  const int ld = mNeigh + 1;
  int iIter, iNeigh, a, b;
  if (!iterIndices) nIter = nAtom;
  for (iIter = 0; iIter < nIter; ++iIter) {
    int iAtom1 = iterIndices ? iterIndices[iIter] : iIter + 1;
    int nOrb1 = nOrbAtom[iAtom1 - 1];
    for (iNeigh = 0; iNeigh <= nNeighbourSK[iAtom1 - 1]; ++iNeigh) {
      int iAtom2f = img2CentCell[iNeighbour[iNeigh + ld * (iAtom1 - 1)] - 1];
      int nOrb2 = nOrbAtom[iAtom2f - 1];
      int iOrig = iPair[iNeigh + ld * (iAtom1 - 1)];
      for (a = 0; a < nOrb1; ++a) {
        double s = 0.0;
        for (b = 0; b < nOrb2; ++b) s += over[iOrig + b + nOrb2 * a] * rho[iOrig + b + nOrb2 * a];
        qq[a + mOrb * (iAtom1 - 1)] += s;
      }
      if (iAtom1 != iAtom2f)
        for (b = 0; b < nOrb2; ++b) {
          double s = 0.0;
          for (a = 0; a < nOrb1; ++a) s += over[iOrig + b + nOrb2 * a] * rho[iOrig + b + nOrb2 * a];
          qq[b + mOrb * (iAtom2f - 1)] += s;
        }
    }
  }
}

/* mullikenPerBlock, populations.F90:306-375.  qq(mOrb, mOrb, nAtom) accumulates. */
void orc_mulliken_per_block(double *qq, int mOrb, const double *over, const double *rho, int nAtom,
                            const int *nOrbAtom, const int *iNeighbour, int mNeigh,
                            const int *nNeighbourSK, const int *img2CentCell, const int *iPair,
                            const int *iterIndices, int nIter)
{
  const int ld = mNeigh + 1;
  int iIter, iNeigh, a, b, c;
  if (!iterIndices) nIter = nAtom;
  for (iIter = 0; iIter < nIter; ++iIter) {
    int iAtom1 = iterIndices ? iterIndices[iIter] : iIter + 1;
    int nOrb1 = nOrbAtom[iAtom1 - 1];
    for (iNeigh = 0; iNeigh <= nNeighbourSK[iAtom1 - 1]; ++iNeigh) {
      int iAtom2f = img2CentCell[iNeighbour[iNeigh + ld * (iAtom1 - 1)] - 1];
      int nOrb2 = nOrbAtom[iAtom2f - 1];
      const double *S = over + iPair[iNeigh + ld * (iAtom1 - 1)];
      const double *P = rho + iPair[iNeigh + ld * (iAtom1 - 1)];
      double *q1 = qq + (size_t)mOrb * mOrb * (iAtom1 - 1);
      /* 0.5 * (rho^T S + S^T rho) : (nOrb1, nOrb1) */
      for (a = 0; a < nOrb1; ++a)
        for (b = 0; b < nOrb1; ++b) {
          double s1 = 0.0, s2 = 0.0;
          for (c = 0; c < nOrb2; ++c) {
            s1 += P[c + nOrb2 * b] * S[c + nOrb2 * a];
            s2 += S[c + nOrb2 * b] * P[c + nOrb2 * a];
          }
          q1[b + mOrb * a] += 0.5 * (s1 + s2);
        }
      if (iAtom1 != iAtom2f) {
        double *q2 = qq + (size_t)mOrb * mOrb * (iAtom2f - 1);
        /* 0.5 * (rho S^T + S rho^T) : (nOrb2, nOrb2) */
        for (a = 0; a < nOrb2; ++a)
          for (b = 0; b < nOrb2; ++b) {
            double s1 = 0.0, s2 = 0.0;
            for (c = 0; c < nOrb1; ++c) {
              s1 += P[b + nOrb2 * c] * S[a + nOrb2 * c];
              s2 += S[b + nOrb2 * c] * P[a + nOrb2 * c];
            }
            q2[b + mOrb * a] += 0.5 * (s1 + s2);
          }
      }
    }
  }
}

/* ------------------------------------------------------------------------------------------
 * dftb/forces.F90 : derivative_shift
 * ---------------------------------------------------------------------------------------- */

/* Per-pair gradient shared by derivative_shift and the stress routines:
 *   derivTmp(ii) = 2 (sum rho.H'_ii - sum Erho.S'_ii)                         (forces.F90:196-199, stress.F90:150-154)
 *                + 2 sum_spin sum 1/2 (S'_ii V1 + V2 S'_ii) . rho_spin          (forces.F90:640-656, stress.F90:291-300) */
static void pairGradient(const orc_model *m, const double *DM, int nSpin, const double *EDM,
                         int sparseSize, const double *shift, int nAtom, const double *coords,
                         const int *species, int iAtom1, int iAtom2, int iAtom2f, int iOrig,
                         int diffType, double delta, double *derivTmp)
{
  const int mo = m->mOrb, sz = mo * mo;
  double hPrime[3 * ORC_MAXORB * ORC_MAXORB], sPrime[3 * ORC_MAXORB * ORC_MAXORB];
  double shiftSprime[ORC_MAXORB * ORC_MAXORB];
  int nOrb1 = m->nOrbSpecies[species[iAtom1 - 1] - 1], nOrb2 = m->nOrbSpecies[species[iAtom2f - 1] - 1];
  int ii, a, b, c, iSpin;
  orc_first_deriv_block(m, 0, hPrime, coords, species, iAtom1, iAtom2, diffType, delta);
  orc_first_deriv_block(m, 1, sPrime, coords, species, iAtom1, iAtom2, diffType, delta);
  for (ii = 0; ii < 3; ++ii) {
    double sH = 0.0, sS = 0.0;
    for (a = 0; a < nOrb1; ++a)
      for (b = 0; b < nOrb2; ++b) {
        sH += DM[iOrig + b + nOrb2 * a] * hPrime[b + mo * a + sz * ii];
        sS += EDM[iOrig + b + nOrb2 * a] * sPrime[b + mo * a + sz * ii];
      }
    derivTmp[ii] = 2.0 * (sH - sS);
  }
  if (shift)
    for (iSpin = 0; iSpin < nSpin; ++iSpin) {
      const double *V1 = shift + (size_t)sz * ((iAtom1 - 1) + (size_t)nAtom * iSpin);
      const double *V2 = shift + (size_t)sz * ((iAtom2f - 1) + (size_t)nAtom * iSpin);
      const double *P = DM + iOrig + (size_t)sparseSize * iSpin;
      for (ii = 0; ii < 3; ++ii) {
        const double *Sp = sPrime + sz * ii;
        double s = 0.0;
        for (a = 0; a < nOrb1; ++a)
          for (b = 0; b < nOrb2; ++b) {
            double x = 0.0, y = 0.0;
            for (c = 0; c < nOrb1; ++c) x += Sp[b + mo * c] * V1[c + mo * a];
            for (c = 0; c < nOrb2; ++c) y += V2[b + mo * c] * Sp[c + mo * a];
            shiftSprime[b + mo * a] = 0.5 * (x + y);
          }
        for (a = 0; a < nOrb1; ++a)
          for (b = 0; b < nOrb2; ++b) s += shiftSprime[b + mo * a] * P[b + nOrb2 * a];
        derivTmp[ii] = derivTmp[ii] + 2.0 * s;
      }
    }
}

/* derivative_shift generic (dftb/forces.F90:27-38): shift == NULL -> derivativeNonSccEuclidian
 * (:119-217), else derivative_blockEuclidean (:548-669).  DM(sparseSize, nSpin) (only the first
 * channel enters the H' term), EDM(sparseSize), shift(mOrb, mOrb, nAtom, nSpin).
 * deriv(3, nAtom) is overwritten.  OpenMP over the outer atom loop with a reduction on deriv,
 * as in the reference (:176, :620). */
void orc_derivative_shift(double *deriv, const orc_model *m, const double *DM, int nSpin,
                          const double *EDM, int sparseSize, const double *shift, int nAtom,
                          const double *coords, const int *species, const int *iNeighbour,
                          int mNeigh, const int *nNeighbourSK, const int *img2CentCell,
                          const int *iPair, int diffType, double delta, const int *iterIndices,
                          int nIter)
{
  const int ld = mNeigh + 1;
  int nThreads = 1, iIter, t, k;
  double *part;
  orc_init();
  if (!iterIndices) nIter = nAtom;
#ifdef _OPENMP
  nThreads = omp_get_max_threads();
#endif
  part = (double *)calloc((size_t)nThreads * 3 * nAtom, sizeof(double));
#pragma omp parallel for schedule(runtime)
  for (iIter = 0; iIter < nIter; ++iIter) {
    double derivTmp[3];
    int iAtom1 = iterIndices ? iterIndices[iIter] : iIter + 1;
    int iNeigh, ii;
    double *my = part;
#ifdef _OPENMP
    my = part + (size_t)omp_get_thread_num() * 3 * nAtom;
#endif
    for (iNeigh = 1; iNeigh <= nNeighbourSK[iAtom1 - 1]; ++iNeigh) {
      int iAtom2 = iNeighbour[iNeigh + ld * (iAtom1 - 1)];
      int iAtom2f = img2CentCell[iAtom2 - 1];
      if (iAtom1 == iAtom2f) continue;
      pairGradient(m, DM, nSpin, EDM, sparseSize, shift, nAtom, coords, species, iAtom1, iAtom2, iAtom2f,
                   iPair[iNeigh + ld * (iAtom1 - 1)], diffType, delta, derivTmp);
      for (ii = 0; ii < 3; ++ii) {
        my[ii + 3 * (iAtom1 - 1)] += derivTmp[ii];
        my[ii + 3 * (iAtom2f - 1)] -= derivTmp[ii];
      }
    }
  }
  for (k = 0; k < 3 * nAtom; ++k) deriv[k] = 0.0;
  for (t = 0; t < nThreads; ++t)
    for (k = 0; k < 3 * nAtom; ++k) deriv[k] += part[(size_t)t * 3 * nAtom + k];
  free(part);
}

/* getNonSCCStress (dftb/stress.F90:70-184; shift == NULL) / getBlockStress (:188-326):
 *   st(jj,ii) += fac * intermed(jj) * vect(ii),  vect = coords(:,iAtom1) - coords(:,iAtom2),
 *   fac = 2 unless iAtom2 is a periodic image of iAtom1 (then 1);  st = -0.5 * st / cellVol.
 * st(3,3) column-major (st[jj + 3*ii]) is overwritten. */
void orc_stress(double *st, const orc_model *m, const double *DM, int nSpin, const double *EDM,
                int sparseSize, const double *shift, int nAtom, const double *coords,
                const int *species, const int *iNeighbour, int mNeigh, const int *nNeighbourSK,
                const int *img2CentCell, const int *iPair, int diffType, double delta, double cellVol,
                const int *iterIndices, int nIter)
{
  const int ld = mNeigh + 1;
  int iIter, k;
  double acc[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  orc_init();
  if (!iterIndices) nIter = nAtom;
#pragma omp parallel for schedule(runtime) reduction(+ : acc[:9])
  for (iIter = 0; iIter < nIter; ++iIter) {
    double intermed[3], vect[3];
    int iAtom1 = iterIndices ? iterIndices[iIter] : iIter + 1;
    int iNeigh, ii, jj;
    for (iNeigh = 1; iNeigh <= nNeighbourSK[iAtom1 - 1]; ++iNeigh) {
      int iAtom2 = iNeighbour[iNeigh + ld * (iAtom1 - 1)];
      int iAtom2f = img2CentCell[iAtom2 - 1];
      double fac = iAtom1 != iAtom2f ? 2.0 : 1.0;
      pairGradient(m, DM, nSpin, EDM, sparseSize, shift, nAtom, coords, species, iAtom1, iAtom2, iAtom2f,
                   iPair[iNeigh + ld * (iAtom1 - 1)], diffType, delta, intermed);
      for (ii = 0; ii < 3; ++ii) vect[ii] = coords[ii + 3 * (iAtom1 - 1)] - coords[ii + 3 * (iAtom2 - 1)];
      for (ii = 0; ii < 3; ++ii)
        for (jj = 0; jj < 3; ++jj) acc[jj + 3 * ii] += fac * intermed[jj] * vect[ii];
    }
  }
  for (k = 0; k < 9; ++k) st[k] = -0.5 * acc[k] / cellVol;
}

/* torchrun exports OMP_NUM_THREADS=1; the CPU baseline must use the cores the reference's OpenMP
 * build would use (OMP_NUM_THREADS = all host cores). */
void orc_set_num_threads(int n)
{
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

int orc_num_threads(void)
{
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* ------------------------------------------------------------------------------------------------
 * Coulomb 1/R matrix of the SCC shifts: invRClusterSerial (dftb/coulomb.F90:704-733),
 * invRPeriodicSerial (:968-1024) with addNeighbourContribsIS (:1027-1053; the dynamic neighbour list
 * of type/dynneighlist.F90:207-260 restated as its loops: central cell iAtom2 > iAtom1, then every
 * other lattice point with iAtom2 >= iAtom1), ewaldReciprocal (:2079-2117), rTerm (:2473-2489), and
 * updateShifts' hemv(..., 'L') (:537-586).  invRMat(ld, nAtom) column-major, lower triangle
 * (row >= column) filled, the rest zero.  norm2 is taken as sqrt(sum(x**2)).
 * rLat(3, nRLat): lattice translations in absolute coordinates, ORIGIN FIRST (getCellTranslations order;
 * the iterator of the dynamic list visits the same points in the same order without the origin).
 * gLat(3, nG): reciprocal lattice points in absolute coordinates, no origin, no inversion pairs. */
void orc_invr_cluster(int nAtom, const double *coord, double *invRMat, int ld)
{
  int i1, i2;
  for (i1 = 0; i1 < nAtom; ++i1) {
    for (i2 = 0; i2 < nAtom; ++i2) invRMat[i2 + (size_t)ld * i1] = 0.0;
    for (i2 = i1 + 1; i2 < nAtom; ++i2) {
      const double x = coord[3 * i1] - coord[3 * i2], y = coord[3 * i1 + 1] - coord[3 * i2 + 1],
                   z = coord[3 * i1 + 2] - coord[3 * i2 + 2];
      invRMat[i2 + (size_t)ld * i1] = 1.0 / sqrt(x * x + y * y + z * z);
    }
  }
}

void orc_invr_periodic(int nAtom, const double *coord, int nRLat, const double *rLat, int nG, const double *gLat,
                       double alpha, double volume, double rCutoff, double *invRMat, int ld)
{
  const double pi = 3.14159265358979323846264338327950288;
  const double cutoff2 = rCutoff * rCutoff;
  int i1;
#pragma omp parallel for schedule(dynamic, 8)
  for (i1 = 0; i1 < nAtom; ++i1) {
    double *col = invRMat + (size_t)ld * i1;
    const double *c1 = coord + 3 * i1;
    int i2, iL, iG;
    for (i2 = 0; i2 < nAtom; ++i2) col[i2] = 0.0;
    /* real space part: central cell (no self term), then all other lattice points */
    for (iL = 0; iL < nRLat; ++iL) {
      const double *L = rLat + 3 * iL;
      for (i2 = (iL == 0 ? i1 + 1 : i1); i2 < nAtom; ++i2) {
        const double nx = coord[3 * i2] + L[0], ny = coord[3 * i2 + 1] + L[1], nz = coord[3 * i2 + 2] + L[2];
        const double dx = c1[0] - nx, dy = c1[1] - ny, dz = c1[2] - nz;
        const double dist2 = dx * dx + dy * dy + dz * dz;
        if (dist2 <= cutoff2) {
          const double rr = sqrt(dist2);
          col[i2] += erfc(alpha * rr) / rr; /* rTerm */
        }
      }
    }
    /* reciprocal space part */
    for (i2 = i1; i2 < nAtom; ++i2) {
      const double rx = c1[0] - coord[3 * i2], ry = c1[1] - coord[3 * i2 + 1], rz = c1[2] - coord[3 * i2 + 2];
      const double g2Factor = -1.0 / (4.0 * alpha * alpha);
      double s = 0.0;
      for (iG = 0; iG < nG; ++iG) {
        const double *g = gLat + 3 * iG;
        const double g2 = g[0] * g[0] + g[1] * g[1] + g[2] * g[2];
        const double gr = g[0] * rx + g[1] * ry + g[2] * rz;
        s += exp(g2Factor * g2) / g2 * cos(gr);
      }
      col[i2] = col[i2] + 2.0 * s * 4.0 * pi / volume - pi / (volume * alpha * alpha);
    }
    /* extra contribution for self interaction */
    col[i1] = col[i1] - 2.0 * alpha / sqrt(pi);
  }
}

/* y = A x with the symmetric A given by its lower triangle (BLAS dsymv 'L' as hemv calls it, beta = 0) */
void orc_symv_lower(int n, const double *A, int ld, const double *x, double *y)
{
  int i, j;
  for (i = 0; i < n; ++i) y[i] = 0.0;
  for (j = 0; j < n; ++j) {
    double t = 0.0;
    y[j] += A[j + (size_t)ld * j] * x[j];
    for (i = j + 1; i < n; ++i) {
      y[i] += A[i + (size_t)ld * j] * x[j];
      t += A[i + (size_t)ld * j] * x[i];
    }
    y[j] += t;
  }
}

/* ------------------------------------------------------------------------------------------------
 * Short-range gamma of the SCC shifts: expGamma / gammaSubExprn_ / expGammaCutoff
 * (dftb/shortgammafuncs.F90:36-148, 470-500), setupShortGammaCutoffs_, updateNrOfNeighbours_,
 * updateShortGammaValues_ (no damping, no H5) and buildShifts_ (dftb/shortgamma.F90:753-936).
 * minHubDiff 0.3125e-5, minHubTol 1e-6, minShortGamma = tolShortGamma = 1e-10, tolSameDist 1e-5
 * (common/accuracy.F90).  Arrays column-major with the reference's index order:
 *   uniqHubbU(mU, nSpecies), shortCutoffs(mU, mU, nSpecies, nSpecies), nNeighU(mU, mU, nSpecies, nAtom),
 *   shortGamma(mU, mU, 0:mNeigh, nAtom), iHubbU(mShell, nSpecies), deltaQShell / shiftShell(mShell, nAtom). */
static const double minHubDiff = 0.3125e-5, tolSameDist = 1.0e-5, tolShortGamma = 1.0e-10, minShortGamma = 1.0e-10;

/* integer powers as gfortran expands them (x**2 = x*x, x**3 = x*x*x, x**4 = (x*x)*(x*x), x**6 = (x*x*x)*(x*x*x)) */
static double pw2(double x) { return x * x; }
static double pw3(double x) { return x * x * x; }
static double pw4(double x) { const double y = x * x; return y * y; }
static double pw6(double x) { const double y = x * x * x; return y * y; }

static double gamma_sub_exprn(double rab, double tau1, double tau2)
{
  return exp(-tau1 * rab) * ((0.5 * pw4(tau2) * tau1 / pw2(pw2(tau1) - pw2(tau2)))
                             - (pw6(tau2) - 3.0 * pw4(tau2) * pw2(tau1)) / (rab * pw3(pw2(tau1) - pw2(tau2))));
}

double orc_exp_gamma(double rab, double Ua, double Ub)
{
  const double tauA = 3.2 * Ua, tauB = 3.2 * Ub;
  if (rab < tolSameDist) {
    if (fabs(Ua - Ub) < minHubDiff) return -0.5 * (Ua + Ub);
    return -0.5 * ((tauA * tauB) / (tauA + tauB) + pw2(tauA * tauB) / pw3(tauA + tauB));
  }
  if (fabs(Ua - Ub) < minHubDiff) {
    const double tauMean = 0.5 * (tauA + tauB);
    return exp(-tauMean * rab) * (1.0 / rab + 0.6875 * tauMean + 0.1875 * rab * pw2(tauMean)
                                  + 0.02083333333333333333 * pw2(rab) * pw3(tauMean));
  }
  return gamma_sub_exprn(rab, tauA, tauB) + gamma_sub_exprn(rab, tauB, tauA);
}

double orc_exp_gamma_cutoff(double U2, double U1)
{
  const double cutValue = minShortGamma;
  double rab = 1.0, minCutOff, maxCutOff, maxGamma, minGamma, lowerGamma, cut, gamma;
  while (orc_exp_gamma(rab, U2, U1) > cutValue) rab = 2.0 * rab;
  if (rab < 2.0) return -1.0; /* "requested tolerance too large" */
  minCutOff = rab + 0.1;
  maxCutOff = 0.5 * rab - 0.1;
  maxGamma = orc_exp_gamma(maxCutOff, U2, U1);
  minGamma = orc_exp_gamma(minCutOff, U2, U1);
  lowerGamma = orc_exp_gamma(minCutOff, U2, U1);
  cut = maxCutOff + 0.1;
  gamma = orc_exp_gamma(cut, U2, U1);
  while ((gamma - lowerGamma) > tolShortGamma) {
    maxCutOff = 0.5 * (cut + minCutOff);
    if ((maxGamma >= minGamma) == (cutValue >= orc_exp_gamma(maxCutOff, U2, U1))) {
      minCutOff = maxCutOff;
      lowerGamma = orc_exp_gamma(minCutOff, U2, U1);
    } else {
      cut = maxCutOff;
      gamma = orc_exp_gamma(cut, U2, U1);
    }
  }
  return minCutOff;
}

/* updateNrOfNeighbours_ + updateShortGammaValues_; maxNeighOut(nAtom) = maxval(nNeigh(:,:,:,iAt)) */
void orc_short_gamma_values(int nAtom, int mNeigh, int mU, int nSpecies, const int *nHubbU, const double *uniqHubbU,
                            const double *shortCutoffs, const int *species, const double *coords,
                            const int *nNeighbour, const int *iNeighbour, const double *neighDist2,
                            int *nNeighU, int *maxNeighOut, double *shortGamma)
{
  const int ld = mNeigh + 1;
  int iAt, iSp, iU1, iU2, iNeigh;
  for (iAt = 0; iAt < nAtom; ++iAt) {
    const int sp1 = species[iAt] - 1;
    int mx = 0;
    for (iSp = 0; iSp < nSpecies; ++iSp)
      for (iU1 = 0; iU1 < mU; ++iU1)
        for (iU2 = 0; iU2 < mU; ++iU2) {
          int v = 0;
          if (iU1 < nHubbU[sp1] && iU2 < nHubbU[iSp]) {
            const double c = shortCutoffs[iU2 + mU * (iU1 + mU * (iSp + nSpecies * sp1))];
            v = orc_search_asc_real_geq(nNeighbour[iAt], neighDist2 + ld * iAt + 1, c * c, tolSameDist2);
          }
          nNeighU[iU2 + mU * (iU1 + mU * (iSp + nSpecies * iAt))] = v;
          if (v > mx) mx = v;
        }
    maxNeighOut[iAt] = mx;
    for (iNeigh = 0; iNeigh < ld * mU * mU; ++iNeigh) shortGamma[(size_t)ld * mU * mU * iAt + iNeigh] = 0.0;
    for (iNeigh = 0; iNeigh <= mx; ++iNeigh) {
      const int iAt2 = iNeighbour[iNeigh + ld * iAt] - 1, sp2 = species[iAt2] - 1;
      const double dx = coords[3 * iAt] - coords[3 * iAt2], dy = coords[3 * iAt + 1] - coords[3 * iAt2 + 1],
                   dz = coords[3 * iAt + 2] - coords[3 * iAt2 + 2];
      const double rab = sqrt(dx * dx + dy * dy + dz * dz);
      for (iU1 = 0; iU1 < nHubbU[sp1]; ++iU1)
        for (iU2 = 0; iU2 < nHubbU[sp2]; ++iU2)
          if (iNeigh <= nNeighU[iU2 + mU * (iU1 + mU * (sp2 + nSpecies * iAt))])
            shortGamma[iU2 + mU * (iU1 + mU * (iNeigh + (size_t)ld * iAt))] =
                orc_exp_gamma(rab, uniqHubbU[iU2 + mU * sp2], uniqHubbU[iU1 + mU * sp1]);
    }
  }
}

/* buildShifts_ (serial branch: all atoms in order) */
void orc_short_gamma_shifts(int nAtom, int mNeigh, int mU, int mShell, int nSpecies, const int *nShell,
                            const int *iHubbU, const int *species, const int *img2CentCell, const int *iNeighbour,
                            const int *maxNeigh, const double *shortGamma, const double *deltaQShell,
                            double *shiftShell)
{
  const int ld = mNeigh + 1;
  int iAt1, iSh1, iSh2, iNeigh;
  (void)nSpecies;
  for (iAt1 = 0; iAt1 < mShell * nAtom; ++iAt1) shiftShell[iAt1] = 0.0;
  for (iAt1 = 0; iAt1 < nAtom; ++iAt1) {
    const int sp1 = species[iAt1] - 1;
    for (iSh1 = 0; iSh1 < nShell[sp1]; ++iSh1) {
      const int iU1 = iHubbU[iSh1 + mShell * sp1] - 1;
      const double dq1 = deltaQShell[iSh1 + mShell * iAt1];
      for (iNeigh = 0; iNeigh <= maxNeigh[iAt1]; ++iNeigh) {
        const int iAt2f = img2CentCell[iNeighbour[iNeigh + ld * iAt1] - 1] - 1, sp2 = species[iAt2f] - 1;
        for (iSh2 = 0; iSh2 < nShell[sp2]; ++iSh2) {
          const int iU2 = iHubbU[iSh2 + mShell * sp2] - 1;
          const double g = shortGamma[iU2 + mU * (iU1 + mU * (iNeigh + (size_t)ld * iAt1))];
          shiftShell[iSh1 + mShell * iAt1] = shiftShell[iSh1 + mShell * iAt1] - g * deltaQShell[iSh2 + mShell * iAt2f];
          if (iAt2f != iAt1)
            shiftShell[iSh2 + mShell * iAt2f] = shiftShell[iSh2 + mShell * iAt2f] - g * dq1;
        }
      }
    }
  }
}

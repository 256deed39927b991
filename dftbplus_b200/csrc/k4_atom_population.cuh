// k4_atom_population.cuh -- the two per-atom population routines of SURVEY 8(a) row a15 that sit on top
// of the per-orbital Mulliken kernel: mullikenPerAtom (dftb/populations.F90:68-112; caller
// getenergies.F90:198, Tr[rho H0] per atom) and getOnsitePopulation (:1195-1222; caller
// dftbplus/main.F90:4726).  One thread per atom, sums in the reference's order (bit exact).
#pragma once
#include "emu_kernel.cuh"

namespace dftbp_b200 {
namespace pop {

// qq(:) = qq + sum(qPerOrbital, dim=1), populations.F90:110.  qPerOrbital(mOrb, nAtom).
NL_KERNEL void pop_atom_sum(int nAtom, int mOrb, const double *__restrict__ qPerOrbital, double *__restrict__ qq) {
  for (long long a = NL_TID; a < nAtom; a += NL_STRIDE) {
    double s = 0.0;
    for (int o = 0; o < mOrb; ++o) s = NL_ADD(s, qPerOrbital[o + (long long)mOrb * a]);
    qq[a] = NL_ADD(qq[a], s);
  }
}

// qq(iAt) = sum(rho(iOrig : iOrig + nOrb*nOrb - 1 : nOrb + 1)), iOrig = iPair(0, iAt) + 1, :1216-1220.
// iPair(0:mNeigh, nAtom) with ld = mNeigh + 1, iAtomStart(nAtom + 1).
NL_KERNEL void pop_onsite(int nAtom, int ld, const int *__restrict__ iPair, const int *__restrict__ iAtomStart,
                          const double *__restrict__ rho, double *__restrict__ qq) {
  for (long long a = NL_TID; a < nAtom; a += NL_STRIDE) {
    const int nOrb = iAtomStart[a + 1] - iAtomStart[a];
    const double *__restrict__ d = rho + iPair[(long long)ld * a];
    double s = 0.0;
    for (int o = 0; o < nOrb; ++o) s = NL_ADD(s, d[(long long)o * (nOrb + 1)]);
    qq[a] = s;
  }
}

}  // namespace pop
}  // namespace dftbp_b200

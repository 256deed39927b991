/*
 * dftbp_b200.h -- C ABI of libdftbp_b200.so: the B200 (sm_100a) implementation of DFTB+'s
 * per-SCC-step data-parallel hot path.
 *
 * This is the drop-in boundary.  Each entry point replaces the body of one Fortran module
 * procedure of the reference (cited below, paths relative to /root/reference/src/dftbp/);
 * INTEGRATION.md shows the ISO_C_BINDING interface block a DFTB+ maintainer adds on the
 * Fortran side.  Plain pointers and sizes only -- no torch / C++ types cross this boundary.
 *
 * Conventions (identical to the reference):
 *   - all reals are IEEE double, complex = interleaved (re, im) doubles (complex(dp));
 *   - arrays are passed exactly as the Fortran caller holds them: column-major, atom and
 *     species VALUES 1-based, iPair (= iSparseStart) offsets 0-based;
 *   - iNeighbour(0:mNeigh, nAtom) and iPair(0:mNeigh, nAtom) arrive as raw (mNeigh+1) x nAtom
 *     int arrays, element [iNeigh + (mNeigh+1)*(iAtom-1)];
 *   - a sparse block is nOrb2 x nOrb1 column-major, rows = orbitals of the NEIGHBOUR.
 *
 * Pointers: every array argument may be a host pointer or a device pointer (detected with
 * cudaPointerGetAttributes).  Host inputs are copied H2D, host outputs D2H, and the call is
 * synchronous for them.  Device pointers are used in place on the context's stream with no
 * synchronisation, so a device-resident driver can chain calls with no PCIe traffic.
 *
 * Errors: every function returns 0 on success, otherwise a DFTBP_B200_E* code; the message is
 * available from dftbp_b200_last_error().  (Reference convention: TStatus + @:RAISE_ERROR,
 * include/error.fypp:87-105, or error() abort, io/message.F90:74-84; the Fortran shim maps a
 * non-zero return onto those.)  There is NO CPU fallback: without a CUDA device
 * dftbp_b200_create fails with DFTBP_B200_ENODEVICE.
 *
 * Threading: a context is used by one host thread at a time (the reference calls these
 * routines from the serial part of the driver); distinct contexts are independent
 * (cf. INSTANCE_SAFE_BUILD, config.cmake:62).
 */
#ifndef DFTBP_B200_H
#define DFTBP_B200_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct dftbp_b200_ctx dftbp_b200_ctx;

enum {
  DFTBP_B200_OK = 0,
  DFTBP_B200_EINVAL = 1,     /* bad argument / call order */
  DFTBP_B200_ENODEVICE = 2,  /* no CUDA device: there is no CPU fallback */
  DFTBP_B200_ECUDA = 3,      /* CUDA runtime / cuSOLVER error */
  DFTBP_B200_ENOTIMPL = 4,   /* valid reference option that is not built (none at present) */
  DFTBP_B200_ENOMEM = 5,
  DFTBP_B200_EGEOMETRY = 6   /* "Atoms i and j too close together" (dftb/periodic.F90:645-648) */
};

/* skIntMethod, dftb/slakoeqgrid.F90:62-65 */
enum { DFTBP_B200_SK_EQGRID_OLD = 1, DFTBP_B200_SK_EQGRID_NEW = 2 };
/* diffTypes, dftb/nonscc.F90:27-38 */
enum { DFTBP_B200_DIFF_FINITE = 1, DFTBP_B200_DIFF_RICHARDSON = 2 };

/* ---- lifecycle ------------------------------------------------------------------------- */

/* One context per TDftbPlus instance / per GPU.  rank / nRanks restate env%mpi%groupComm
 * (common/schedule.F90:51-76): with nRanks > 1 the pair-parallel routines (build_*, mulliken_*,
 * derivative_shift) only process this rank's share of the atoms and leave the rest of the
 * output zero, exactly like the reference before assembleChunks (schedule.F90:170-182); the
 * caller sums over ranks (MPI_Allreduce in the reference, NCCL in the host layer here). */
int dftbp_b200_create(dftbp_b200_ctx **ctx, int device, int rank, int nRanks);
int dftbp_b200_destroy(dftbp_b200_ctx *ctx);
const char *dftbp_b200_last_error(const dftbp_b200_ctx *ctx);
const char *dftbp_b200_version(void);

/* Use an existing CUDA stream (cudaStream_t passed as void*); NULL = the context's own
 * (non-blocking) stream; (void*)1 = cudaStreamLegacy, the default stream most host frameworks
 * (PyTorch included) enqueue on.  Device-pointer callers must put the library on the stream that
 * produces / consumes their buffers.  dftbp_b200_synchronize blocks until all work queued by this
 * context has finished. */
int dftbp_b200_set_stream(dftbp_b200_ctx *ctx, void *cuda_stream);
void *dftbp_b200_get_stream(dftbp_b200_ctx *ctx);
int dftbp_b200_synchronize(dftbp_b200_ctx *ctx);

/* ---- once per run: flattened TOrbitals and TSlakoCont ----------------------------------- */

/* TOrbitals, type/orbitals.F90:20-51.  angShell is (mShell, nSpecies) column-major; angular
 * momenta 0..3 (mAngRot_, dftb/sk.F90:28).  nOrbSpecies is derived. */
int dftbp_b200_set_orbitals(dftbp_b200_ctx *ctx, int nSpecies, int mShell, const int *nShell,
                            const int *angShell);

/* SlakoCont_addTableEqGrid + SlakoEqGrid_init, dftb/slakocont.F90:97-111 and
 * dftb/slakoeqgrid.F90:95-121.  which: 0 = Hamiltonian container, 1 = overlap container.
 * (sp1, sp2) (1-based) is the pair the reference later fetches with
 * getSKIntegrals(skCont, sk, dist, sp1, sp2), sp1 = species of the atom, sp2 = species of its
 * neighbour (slakocont.F90:156-176).  skTab is (nGrid, nInteg) column-major, row j <-> r = j*dist;
 * nInteg must equal sum over shell pairs (sp1 outer, sp2 inner) of min(l1,l2)+1
 * (dftbplus/parser.F90:4061-4130). */
int dftbp_b200_set_sk_table(dftbp_b200_ctx *ctx, int which, int sp1, int sp2, int nGrid,
                            int nInteg, double dist, const double *skTab, int skIntMethod);

/* SlakoCont_getCutoff, slakocont.F90:139-150 (max over tables of nGrid*dist + distFudge) */
int dftbp_b200_get_sk_cutoff(dftbp_b200_ctx *ctx, double *cutoff);

/* ---- once per geometry (handleCoordinateChange, dftbplus/main.F90:2413-2577) ------------- */

/* coords(3, nAllAtom), species(nAllAtom), img2CentCell(nAllAtom), iCellVec(nAllAtom),
 * cellVec(3, nCellVec), nNeighbourSK(nAtom), iNeighbour(0:mNeigh, nAtom), iPair(0:mNeigh, nAtom),
 * iAtomStart(nAtom+1).  All are copied to the device. */
int dftbp_b200_set_geometry(dftbp_b200_ctx *ctx, int nAtom, int nAllAtom, int mNeigh,
                            const double *coords, const int *species, const int *img2CentCell,
                            const int *iCellVec, int nCellVec, const double *cellVec,
                            const int *nNeighbourSK, const int *iNeighbour, const int *iPair,
                            const int *iAtomStart, int sparseSize);

/* ---- N2: neighbour list and sparse descriptor on the device -------------------- */

/* getCellTranslations, dftb/periodic.F90:227-263 (host side; runs when the lattice changes):
 * the zero vector first, then every other point of the parallelepiped +-(floor(cutoff*|b_i|)+1),
 * third index fastest (type/latpointiter.F90:214-274).  latVec(3,3), recVec2p(3,3) column-major
 * (reciprocal vectors in units of 2 pi).  cellVec / rCellVec (3, maxCellVec) may be NULL to query
 * *nCellVec. */
int dftbp_b200_cell_translations(const double *latVec, const double *recVec2p, double cutoff,
                                 int maxCellVec, int *nCellVec, double *cellVec, double *rCellVec);

/* foldCoordsToCell, pbc3d branch without contacts (dftb/boundarycond.F90:369-380; the step before
 * the list in handleCoordinateChange, dftbplus/main.F90:2559-2561): coord(3, nAtom) is folded into
 * the cell of latVec(3,3) (column-major, column i = i-th lattice vector) in place, with the
 * reference's invert33 and operation order.  Cluster geometries are not folded (no call). */
int dftbp_b200_fold_coords_to_cell(dftbp_b200_ctx *ctx, int nAtom, double *coord, const double *latVec);

/* updateNeighbourListAndSpecies (dftb/periodic.F90:380-440, 446-694; no helical boundary
 * conditions) + getNrOfNeighboursForAll(skCutoff) (:996-1054) + getSparseDescriptor
 * (dftb/sparse2dense.F90:3617-3662), as called by handleCoordinateChange
 * (dftbplus/main.F90:2562-2578), with a cell list on the device instead of the reference's
 * O(nCellVec nAtom^2) scan.  coord0(3, nAtom) (already folded into the cell by the caller),
 * species0(nAtom), cellVec / rCellVec(3, nCellVec) with the zero vector first (nCellVec = 1 and
 * zeros for a cluster), cutoff = list cutoff (mCutOff), skCutoff <= cutoff.  Image numbering, row
 * order (including the order inside shells of equal distance, which fixes the iPair layout) and
 * all values are the reference's.  The result stays on the device; the sizes the caller needs to
 * allocate its copies come back in nAllAtom, mNeigh (= maxval(nNeighbour)) and sparseSize.
 * symmetric = 0: the non-symmetric list every routine of this library expects; != 0: the reference's
 * `symmetric` option (iAtom2 runs over all atoms, :580-584; symNeighbourList of the hybrid functionals,
 * dftbplus/main.F90:2587) for the host's use -- such a list cannot be adopted.
 * Atoms closer than minNeighDist: DFTBP_B200_EGEOMETRY, message as the reference's. */
int dftbp_b200_update_neighbour_list(dftbp_b200_ctx *ctx, int nAtom, const double *coord0,
                                     const int *species0, int nCellVec, const double *cellVec,
                                     const double *rCellVec, double cutoff, double skCutoff,
                                     int symmetric, int *nAllAtom, int *mNeigh, int *sparseSize);

/* Copies of the arrays of the last update (any pointer may be NULL; host or device memory):
 * coords(3, nAllAtom), species / img2CentCell / iCellVec(nAllAtom), nNeighbour(nAtom),
 * iNeighbour / neighDist2 / iPair(0:mNeigh, nAtom), nNeighbourSK(nAtom), iAtomStart(nAtom+1). */
int dftbp_b200_get_neighbour_list(dftbp_b200_ctx *ctx, double *coords, int *species, int *img2CentCell,
                                  int *iCellVec, int *nNeighbour, int *iNeighbour, double *neighDist2,
                                  int *nNeighbourSK, int *iPair, int *iAtomStart);

/* getNrOfNeighboursForAll, dftb/periodic.F90:996-1054 (search_asc_real_geq with tolSameDist2), for
 * another cutoff <= the list cutoff on the list of the last update (the reference counts again for
 * the repulsive / CAM ranges, e.g. dftbplus/main.F90:2583).  nNeighbour(nAtom) out. */
int dftbp_b200_nr_of_neighbours(dftbp_b200_ctx *ctx, double cutoff, int *nNeighbour);

/* Make the list of the last update the context's geometry (what dftbp_b200_set_geometry does
 * with caller arrays): the arrays are taken from the device copies, the caller passes nothing.
 * (The work lists / block descriptors of set_geometry are still formed on the host from a read-back
 * of the index arrays.) */
int dftbp_b200_adopt_neighbour_list(dftbp_b200_ctx *ctx);

/* Coordinates only (same neighbour map): geometry steps that keep the list, and finite
 * differences.  coords(3, nAllAtom). */
int dftbp_b200_update_coords(dftbp_b200_ctx *ctx, const double *coords);

/* ---- K1: Slater-Koster pair build -------------------------------------------------------- */

/* buildH0, dftb/nonscc.F90:66-124.  selfegy(mShell, nSpecies); ham(sparseSize) out. */
int dftbp_b200_build_h0(dftbp_b200_ctx *ctx, const double *selfegy, double *ham);
/* buildS, dftb/nonscc.F90:128-181.  over(sparseSize) out. */
int dftbp_b200_build_s(dftbp_b200_ctx *ctx, double *over);
/* both in one launch (the reference calls them back to back, dftbplus/main.F90:1219-1221) */
int dftbp_b200_build_h0_s(dftbp_b200_ctx *ctx, const double *selfegy, double *ham, double *over);

/* ---- K2: sparse <-> dense ----------------------------------------------------------------- */

/* unpackHS_real, dftb/sparse2dense.F90:266-321.  square(ld, nOrb), ld >= nOrb; zeroed, then
 * off-site blocks go to the lower triangle. */
int dftbp_b200_unpack_hs_real(dftbp_b200_ctx *ctx, const double *orig, double *square, int ld);
/* unpackHS_cmplx_kpts, dftb/sparse2dense.F90:185-260.  kPoint in relative coordinates. */
int dftbp_b200_unpack_hs_cplx(dftbp_b200_ctx *ctx, const double *orig, const double kPoint[3],
                              void *square_c128, int ld);
/* packHS_real, dftb/sparse2dense.F90:1024-1096.  primitive(sparseSize) ACCUMULATES. */
int dftbp_b200_pack_hs_real(dftbp_b200_ctx *ctx, double *primitive, const double *square, int ld);
/* packHS_cmplx_kpts, dftb/sparse2dense.F90:820-918. */
int dftbp_b200_pack_hs_cplx(dftbp_b200_ctx *ctx, double *primitive, const void *square_c128,
                            int ld, const double kPoint[3], double kWeight);

/* ---- K3: density matrix -------------------------------------------------------------------- */

/* TDensityMatrix%getDensityMatrix / getEDensityMatrix -> fullDensityMatrix_real and
 * fullEnergyDensityMatrix_real, dftb/densitymatrix.F90:123-233,530-591,664-733:
 *   dm = lower triangle of  C(:,1:nLevels) diag(w(1:nLevels)) C(:,1:nLevels)^T,
 * nLevels = index of the last |filling| >= epsilon(1.0).  filling(n) decides nLevels;
 * weights(n) are the weights (== filling for the density matrix, filling*eigen for the
 * energy-weighted one); pass weights = NULL to use filling.  Weights of either sign are handled
 * directly (the reference's sqrt / shift-by-arbitraryConstant reformulations exist only because
 * BLAS xSYRK needs a positive-definite form).  dm(ld, n): the strict upper triangle is zeroed,
 * eigvecs(ld, n) is not modified. */
int dftbp_b200_density_matrix_real(dftbp_b200_ctx *ctx, double *dm, const double *eigvecs, int n,
                                   int ld, const double *filling, const double *weights);
/* fullDensityMatrix_cmplx / fullEnergyDensityMatrix_cmplx, densitymatrix.F90:597-658,739-806:
 * dm = lower triangle of C diag(w) C^H. */
int dftbp_b200_density_matrix_cplx(dftbp_b200_ctx *ctx, void *dm_c128, const void *eigvecs_c128,
                                   int n, int ld, const double *filling, const double *weights);

/* Multi-GPU split of ONE density matrix (more GPUs than (k-point, spin) pairs): only
 * the lower-triangle tiles that overlap the dense columns colFirst..colLast (1-based, inclusive)
 * are computed; entries outside the panel are unspecified.  With panels that start and end on atom
 * boundaries, dftbp_b200_pack_hs_*_atoms folds exactly the panel's blocks (blocks of iAtom1 in
 * atomFirst..atomLast, 1-based) into the sparse array, and the sum of the ranks' sparse arrays
 * (the all-reduce of rhoPrim the reference does anyway, main.F90:4179) is the full result. */
int dftbp_b200_density_matrix_real_cols(dftbp_b200_ctx *ctx, double *dm, const double *eigvecs, int n,
                                        int ld, const double *filling, const double *weights,
                                        int colFirst, int colLast);
int dftbp_b200_density_matrix_cplx_cols(dftbp_b200_ctx *ctx, void *dm_c128, const void *eigvecs_c128,
                                        int n, int ld, const double *filling, const double *weights,
                                        int colFirst, int colLast);
int dftbp_b200_pack_hs_real_atoms(dftbp_b200_ctx *ctx, double *primitive, const double *square, int ld,
                                  int atomFirst, int atomLast);
int dftbp_b200_pack_hs_cplx_atoms(dftbp_b200_ctx *ctx, double *primitive, const void *square_c128,
                                  int ld, const double kPoint[3], double kWeight, int atomFirst,
                                  int atomLast);

/* ---- K4: Mulliken populations -------------------------------------------------------------- */

/* mullikenPerOrbital, dftb/populations.F90:120-187.  qq(mOrb, nAtom) ACCUMULATES. */
int dftbp_b200_mulliken_orbital(dftbp_b200_ctx *ctx, double *qq, const double *over,
                                const double *rho);
/* mullikenPerBlock, dftb/populations.F90:306-375.  qq(mOrb, mOrb, nAtom) ACCUMULATES. */
int dftbp_b200_mulliken_block(dftbp_b200_ctx *ctx, double *qq, const double *over,
                              const double *rho);
/* mullikenPerAtom, dftb/populations.F90:68-112 (e.g. Tr[rho H0] per atom, dftb/getenergies.F90:198):
 * qq(nAtom) ACCUMULATES sum over the orbitals of mullikenPerOrbital. */
int dftbp_b200_mulliken_atom(dftbp_b200_ctx *ctx, double *qq, const double *over, const double *rho);
/* getOnsitePopulation, dftb/populations.F90:1195-1222 (caller dftbplus/main.F90:4726): qq(nAtom) =
 * trace of the on-site block of rho (net atomic population); OVERWRITTEN. */
int dftbp_b200_onsite_population(dftbp_b200_ctx *ctx, const double *rho, double *qq);

/* getNonSCCStress (shift == NULL) / getBlockStress, dftb/stress.F90:70-184 / 188-326 (SURVEY 8(f)
 * N3; callers dftbplus/main.F90:1829, 7491): the pair loop of derivative_shift with the outer
 * product of the pair gradient and the bond vector, st = -0.5 / cellVol * sum fac * g(jj) * vect(ii),
 * fac = 2 (1 for an atom and its own periodic image).  st(3,3) column-major, OVERWRITTEN; the other
 * arguments as for dftbp_b200_derivative_shift.  With nRanks > 1: this rank's atoms only. */
int dftbp_b200_stress(dftbp_b200_ctx *ctx, double *st, const double *DM, int nSpin, const double *EDM,
                      const double *shift, int diffType, double delta, double cellVol);

/* ---- K6: SCC contribution to the sparse Hamiltonian ------------------------ */

/* addShift_atom, dftb/shift.F90:41-127 (caller: getSccHamiltonian, dftbplus/main.F90:1449):
 * ham(:, iSpin) += over * 0.5 * (shift(iAt1, iSpin) + shift(iAt2f, iSpin)) for every sparse block.
 * ham(sparseSize, nSpin) in/out, over(sparseSize), shift(nAtom, nSpin), all column-major.
 * isInputZero != 0: ham is zeroed first and only this rank's workload-balanced atoms are added
 * (the caller sums the ranks, assembleChunks); == 0: every rank adds all atoms, like the reference. */
int dftbp_b200_add_shift_atom(dftbp_b200_ctx *ctx, double *ham, const double *over,
                              const double *shift, int nSpin, int isInputZero);
/* addShift_block, dftb/shift.F90:132-228: ham(block, iSpin) += 0.5 * (S . shift(:,:,iAt1,iSpin) +
 * shift(:,:,iAt2f,iSpin) . S); shift(mOrb, mOrb, nAtom, nSpin). */
int dftbp_b200_add_shift_block(dftbp_b200_ctx *ctx, double *ham, const double *over,
                               const double *shift, int nSpin, int isInputZero);

/* ---- K5: force contraction ----------------------------------------------------------------- */

/* Optional finer decomposition of derivative_shift / stress across GPUs than the reference's per-atom
 * one (getIndicesWithWorkload, common/schedule.F90:267-307): subsequent force calls of this context
 * take the pairs part, part + nParts, ... (0-based, counted along the neighbour list) of the geometry --
 * perfectly balanced, and every GPU can keep ONE context (nRanks = 1: full H0 / S build) instead of a
 * second, atom-sharded one.  The share is a property of the pair (its position in the neighbour list),
 * not of the device-built schedule, whose order inside a distance bin differs from run to run and from
 * GPU to GPU: contexts on different GPUs holding the same geometry take disjoint, complete shares.
 * The partial deriv / stress arrays add up to the full result, i.e. the caller's assembleChunks
 * (all-reduce) is unchanged.  Default: part 0 of 1. */
int dftbp_b200_set_pair_partition(dftbp_b200_ctx *ctx, int part, int nParts);

/* derivative_shift generic, dftb/forces.F90:27-38: shift == NULL -> derivativeNonSccEuclidian
 * (:119-217), else derivative_blockEuclidean (:548-669).  DM(sparseSize, nSpin),
 * EDM(sparseSize), shift(mOrb, mOrb, nAtom, nSpin), deriv(3, nAtom) is overwritten.
 * diffType / delta restate TNonSccDiff (dftb/nonscc.F90:39-60); delta <= 0 selects the
 * reference default epsilon(1.0)**0.25. */
int dftbp_b200_derivative_shift(dftbp_b200_ctx *ctx, double *deriv, const double *DM, int nSpin,
                                const double *EDM, const double *shift, int diffType,
                                double delta);

/* ---- N4 (first piece): electron filling on the device ------------------------------------- */

/* fillingTypes, dftb/etemp.F90:28-40 */
enum { DFTBP_B200_FILLING_FERMI = 0, DFTBP_B200_FILLING_GAUSSIAN = 1, DFTBP_B200_FILLING_METHFESSEL = 2 };

/* Efilling, dftb/etemp.F90:59-191 (caller getFillingsAndBandEnergies, dftbplus/main.F90:4597-4648; SURVEY
 * 8(f) N4): common Fermi level Ef for nElectrons over eigenvals(nLevels, nKpt, nSpin) with k-point weights
 * kWeight(nKpt) at thermal energy kT, fillings (one electron per level: the caller halves nElectrons and
 * doubles the results for spin-unpolarised runs, as the reference does), band energy Ebs(nSpin), entropy
 * term TS(nSpin), energy extrapolated to 0 K E0(nSpin).  Same control flow as the reference (mid-gap
 * shortcut for integer electron numbers, bracket widening, bisection to elecTol, cut at elecTol, final
 * rescaling); distrib = Fermi only (Gaussian / Methfessel-Paxton return DFTBP_B200_ENOTIMPL).  Sits between
 * dftbp_b200_hegv_* and dftbp_b200_density_matrix_*: with device pointers the eigenvalues and fillings
 * never leave the GPU.  "Fermi level search did not converge." is returned as DFTBP_B200_EINVAL when the
 * small outputs are host pointers. */
int dftbp_b200_efilling(dftbp_b200_ctx *ctx, double *Ebs, double *Ef, double *TS, double *E0,
                        double *filling, const double *eigenvals, int nLevels, int nKpt, int nSpin,
                        double nElectrons, double kT, const double *kWeight, int distrib);

/* The Coulomb 1/R matrix of the SCC shifts (second piece).  invRMat(ld, nAtom) column-major,
 * lower triangle (row >= column) filled, everything else zero -- the storage hemv(..., 'L') reads.
 * coord(3, nAtom): the central-cell atoms.  Host or device pointers throughout.
 *
 * invr_cluster:   invRClusterSerial, dftb/coulomb.F90:704-733 (caller TCoulomb%updateCoords, :291-327).
 * invr_periodic:  invRPeriodicSerial, dftb/coulomb.F90:968-1024: the Ewald sum with parameter alpha over
 *   rLat(3, nRLat) -- the lattice translations in absolute coordinates, origin first, as
 *   getCellTranslations(.., maxREwald) returns them (this%rLatPoints_; the dynamic neighbour list of the
 *   reference, type/dynneighlist.F90, visits the same points in the same order), real-space terms kept for
 *   dist^2 <= rCutoff^2 -- and gLat(3, nG), the reciprocal points this%gLatPoints_ (absolute, no origin,
 *   one of every inversion pair); volume = cell volume.  The reciprocal part is evaluated as a weighted Gram
 *   matrix of structure factors on the FP64 tensor cores (same value up to rounding).
 * coulomb_shift:  TCoulomb%updateShifts, dftb/coulomb.F90:537-586: shiftPerAtom(:) = hemv(invRMat,
 *   deltaQAtom, 'L') (overwrites shiftPerAtom). */
int dftbp_b200_invr_cluster(dftbp_b200_ctx *ctx, int nAtom, const double *coord, double *invRMat, int ld);
int dftbp_b200_invr_periodic(dftbp_b200_ctx *ctx, int nAtom, const double *coord, int nRLat, const double *rLat,
                             int nG, const double *gLat, double alpha, double volume, double rCutoff,
                             double *invRMat, int ld);
int dftbp_b200_coulomb_shift(dftbp_b200_ctx *ctx, int nAtom, const double *invRMat, int ld,
                             const double *deltaQAtom, double *shiftPerAtom);

/* Short-range gamma of the SCC shifts (third piece), on the device neighbour list of the last
 * dftbp_b200_update_neighbour_list (its cutoff must cover the largest short-range cutoff, as mCutOff does in
 * the reference).  Arrays in the reference's layout, host or device pointers.
 *
 * short_gamma_update: TShortGamma%updateCoords, dftb/shortgamma.F90:177-196 = updateNrOfNeighbours_ (:779-801,
 *   getNrOfNeighbours per (iU2, iU1, species, atom) with shortCutoffs(mHubbU, mHubbU, nSpecies, nSpecies) from
 *   setupShortGammaCutoffs_) + updateShortGammaValues_ (:804-857, expGamma of shortgammafuncs.F90:97-148; no
 *   damping, no H5 correction -- those stay with the Fortran routine).  nHubbU(nSpecies), uniqHubbU(mHubbU,
 *   nSpecies): this%hubbU_ (TUniqueHubbard).  mHubbU <= 4.  Call it after every neighbour-list update.
 * short_gamma_get: nNeigh_(mHubbU, mHubbU, nSpecies, nAtom), maxval over the first three dimensions per atom,
 *   shortGamma_(mHubbU, mHubbU, 0:mNeigh, nAtom) of the last update (any pointer may be NULL).
 * short_gamma_shifts: TShortGamma%updateShifts -> buildShifts_, dftb/shortgamma.F90:860-933: shiftShell(mShell,
 *   nAtom) from deltaQShell(mShell, nAtom) and iHubbU(mShell, nSpecies) (overwrites shiftShell). */
int dftbp_b200_short_gamma_update(dftbp_b200_ctx *ctx, int nSpecies, int mHubbU, const int *nHubbU,
                                  const double *uniqHubbU, const double *shortCutoffs);
int dftbp_b200_short_gamma_get(dftbp_b200_ctx *ctx, int *nNeighU, int *maxNeigh, double *shortGamma);
int dftbp_b200_short_gamma_shifts(dftbp_b200_ctx *ctx, const int *iHubbU, const double *deltaQShell,
                                  double *shiftShell);

/* The charge mixer of the SCC loop with its vectors on the device: the modified Broyden mixer
 * TBroydenMixerReal, mixer/broydenmixer.F90 -- init (:110-139; parser defaults mixParam 0.2, omega0 0.01,
 * minWeight 1, maxWeight 1e5, weightFac 1e-2: dftbplus/parser.F90:5786-5790), reset (:143-170, once per SCC
 * cycle), mix (:180-342): qInpResult(nElem) holds the input charges on entry and the mixed charges on exit,
 * qDiff(nElem) = output - input charges.  Host or device pointers; with device pointers the charges go from the
 * Mulliken kernel to the shift kernels without leaving the GPU (the call itself reads back 2 + 2 (iter - 2) + 1
 * scalar products and solves the small system on the host, as the reference does with LAPACK).  More than
 * maxSccIter calls after a reset fail with "Broyden mixer: Maximal nr. of steps exceeded". */
int dftbp_b200_mixer_init(dftbp_b200_ctx *ctx, int maxSccIter, double mixParam, double omega0, double minWeight,
                          double maxWeight, double weightFac);
int dftbp_b200_mixer_reset(dftbp_b200_ctx *ctx, int nElem);
int dftbp_b200_mixer_mix(dftbp_b200_ctx *ctx, double *qInpResult, const double *qDiff, int nElem);

/* ---- not rewritten: dense generalized eigensolve via cuSOLVER ------------------------------- */

/* diagDenseMtx, dftbplus/eigenvects.F90:85-100 (hegvd branch): H C = S C diag(eigen), lower
 * triangles referenced; on exit H holds the eigenvectors, S its Cholesky factor.  jobz 'V'/'N'.
 * H, S (ld, n) and eigen(n) may be host or device pointers. */
int dftbp_b200_hegv_real(dftbp_b200_ctx *ctx, double *H, double *S, double *eigen, int n, int ld,
                         char jobz);
int dftbp_b200_hegv_cplx(dftbp_b200_ctx *ctx, void *H_c128, void *S_c128, double *eigen, int n,
                         int ld, char jobz);

/* ---- introspection (bench / tests) --------------------------------------------------------- */

/* Device-time accounting per routine with CUDA events on the context's stream; slots are named
 * after the reference's timers (common/environment.F90:96-134). */
enum {
  DFTBP_B200_TIMER_SPARSE_H0S = 0,     /* "Sparse H0 and S build"   */
  DFTBP_B200_TIMER_SPARSE_TO_DENSE,    /* "Sparse to dense"         */
  DFTBP_B200_TIMER_DENSE_TO_SPARSE,    /* "Dense to sparse"         */
  DFTBP_B200_TIMER_DENSITY_MATRIX,     /* "Density matrix creation" */
  DFTBP_B200_TIMER_MULLIKEN,
  DFTBP_B200_TIMER_FORCE_CALC,         /* "Force calculation"       */
  DFTBP_B200_TIMER_DIAGONALIZATION,    /* "Diagonalisation"         */
  DFTBP_B200_TIMER_ADD_SHIFT,          /* addShift (inside "SCC")   */
  DFTBP_B200_TIMER_NEIGHBOUR_LIST,     /* updateNeighbourList + getSparseDescriptor */
  DFTBP_B200_TIMER_COUNT
};
int dftbp_b200_timer_enable(dftbp_b200_ctx *ctx, int on);   /* also resets the accumulators */
/* synchronises the stream; total device milliseconds and number of calls of one slot */
int dftbp_b200_timer_read(dftbp_b200_ctx *ctx, int slot, double *ms, long long *calls);

/* number of kernels this context has launched since creation */
long long dftbp_b200_kernel_launches(const dftbp_b200_ctx *ctx);
/* number of (iAtom, iNeigh >= 1) pairs in this rank's share of the current geometry */
long long dftbp_b200_num_pairs(const dftbp_b200_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif /* DFTBP_B200_H */

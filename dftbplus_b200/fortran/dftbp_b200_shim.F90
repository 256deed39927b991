!> ISO_C_BINDING shim: lets the DFTB+ host call libdftbp_b200.so behind its own routine signatures.
!>
!> Delivered as source only -- this image has no Fortran compiler, so the file has been reviewed
!> against the reference signatures but not compiled.  The wrapper routines take exactly the dummy
!> arguments of the reference routines they replace (cited per routine), minus the derived types
!> that the one-off set_* calls have already flattened into the context.
module dftbp_b200_shim
  use, intrinsic :: iso_c_binding
  implicit none
  private

  public :: TB200Ctx, b200_init, b200_destroy, b200_setGeometry, b200_updateCoords
  public :: b200_updateNeighbourListAndSpecies, b200_foldCoordsToCell, b200_getNrOfNeighboursForAll
  public :: b200_buildH0, b200_buildS, b200_unpackHS_real, b200_unpackHS_cmplx
  public :: b200_packHS_real, b200_packHS_cmplx
  public :: b200_getDensityMatrix_real, b200_getDensityMatrix_cmplx
  public :: b200_getEDensityMatrix_real, b200_getEDensityMatrix_cmplx
  public :: b200_mullikenPerOrbital, b200_mullikenPerBlock, b200_mullikenPerAtom, b200_getOnsitePopulation
  public :: b200_derivative_shift, b200_addShift_block, b200_addShift_atom, b200_getBlockStress
  public :: b200_hegv_real, b200_hegv_cmplx, b200_Efilling
  public :: b200_invRCluster, b200_invRPeriodic, b200_coulombShifts
  public :: b200_shortGammaUpdateCoords, b200_shortGammaShifts
  public :: b200_BroydenMixer_init, b200_BroydenMixer_reset, b200_BroydenMixer_mix

  integer, parameter :: dp = c_double

  type :: TB200Ctx
    type(c_ptr) :: h = c_null_ptr
  end type TB200Ctx

  interface
    integer(c_int) function dftbp_b200_create(ctx, device, rank, nRanks) bind(C)
      import :: c_ptr, c_int
      type(c_ptr), intent(out) :: ctx
      integer(c_int), value :: device, rank, nRanks
    end function
    function dftbp_b200_last_error(ctx) bind(C) result(msg)
      import :: c_ptr
      type(c_ptr), value :: ctx
      type(c_ptr) :: msg
    end function
    integer(c_int) function dftbp_b200_set_orbitals(ctx, nSpecies, mShell, nShell, angShell) bind(C)
      import :: c_ptr, c_int
      type(c_ptr), value :: ctx
      integer(c_int), value :: nSpecies, mShell
      integer(c_int), intent(in) :: nShell(*), angShell(*)
    end function
    integer(c_int) function dftbp_b200_set_sk_table(ctx, which, sp1, sp2, nGrid, nInteg, dist, skTab,&
        & skIntMethod) bind(C)
      import :: c_ptr, c_int, c_double
      type(c_ptr), value :: ctx
      integer(c_int), value :: which, sp1, sp2, nGrid, nInteg, skIntMethod
      real(c_double), value :: dist
      real(c_double), intent(in) :: skTab(*)
    end function
    integer(c_int) function dftbp_b200_set_geometry(ctx, nAtom, nAllAtom, mNeigh, coords, species,&
        & img2CentCell, iCellVec, nCellVec, cellVec, nNeighbourSK, iNeighbour, iPair, iAtomStart,&
        & sparseSize) bind(C)
      import :: c_ptr, c_int, c_double
      type(c_ptr), value :: ctx
      integer(c_int), value :: nAtom, nAllAtom, mNeigh, nCellVec, sparseSize
      real(c_double), intent(in) :: coords(*), cellVec(*)
      integer(c_int), intent(in) :: species(*), img2CentCell(*), iCellVec(*), nNeighbourSK(*)
      integer(c_int), intent(in) :: iNeighbour(*), iPair(*), iAtomStart(*)
    end function
    integer(c_int) function dftbp_b200_build_h0(ctx, selfegy, ham) bind(C)
      import :: c_ptr, c_int, c_double
      type(c_ptr), value :: ctx
      real(c_double), intent(in) :: selfegy(*)
      real(c_double), intent(out) :: ham(*)
    end function
    integer(c_int) function dftbp_b200_build_s(ctx, over) bind(C)
      import :: c_ptr, c_int, c_double
      type(c_ptr), value :: ctx
      real(c_double), intent(out) :: over(*)
    end function
    integer(c_int) function dftbp_b200_unpack_hs_real(ctx, orig, square, ld) bind(C)
      import :: c_ptr, c_int, c_double
      type(c_ptr), value :: ctx
      real(c_double), intent(in) :: orig(*)
      real(c_double), intent(out) :: square(*)
      integer(c_int), value :: ld
    end function
    integer(c_int) function dftbp_b200_unpack_hs_cplx(ctx, orig, kPoint, square, ld) bind(C)
      import :: c_ptr, c_int, c_double, c_double_complex
      type(c_ptr), value :: ctx
      real(c_double), intent(in) :: orig(*), kPoint(3)
      complex(c_double_complex), intent(out) :: square(*)
      integer(c_int), value :: ld
    end function
    integer(c_int) function dftbp_b200_pack_hs_real(ctx, primitive, square, ld) bind(C)
      import :: c_ptr, c_int, c_double
      type(c_ptr), value :: ctx
      real(c_double), intent(inout) :: primitive(*)
      real(c_double), intent(in) :: square(*)
      integer(c_int), value :: ld
    end function
    integer(c_int) function dftbp_b200_pack_hs_cplx(ctx, primitive, square, ld, kPoint, kWeight) bind(C)
      import :: c_ptr, c_int, c_double, c_double_complex
      type(c_ptr), value :: ctx
      real(c_double), intent(inout) :: primitive(*)
      complex(c_double_complex), intent(in) :: square(*)
      integer(c_int), value :: ld
      real(c_double), intent(in) :: kPoint(3)
      real(c_double), value :: kWeight
    end function
    integer(c_int) function dftbp_b200_density_matrix_real(ctx, dm, eigvecs, n, ld, filling, weights)&
        & bind(C)
      import :: c_ptr, c_int, c_double
      type(c_ptr), value :: ctx
      real(c_double), intent(out) :: dm(*)
      real(c_double), intent(in) :: eigvecs(*), filling(*)
      integer(c_int), value :: n, ld
      type(c_ptr), value :: weights
    end function
    integer(c_int) function dftbp_b200_mulliken_orbital(ctx, qq, over, rho) bind(C)
      import :: c_ptr, c_int, c_double
      type(c_ptr), value :: ctx
      real(c_double), intent(inout) :: qq(*)
      real(c_double), intent(in) :: over(*), rho(*)
    end function
    integer(c_int) function dftbp_b200_stress(ctx, st, DM, nSpin, EDM, shift, diffType, delta, cellVol)&
        & bind(C)
      import :: c_ptr, c_int, c_double
      type(c_ptr), value :: ctx
      real(c_double), intent(out) :: st(3,3)
      real(c_double), intent(in) :: DM(*), EDM(*)
      integer(c_int), value :: nSpin, diffType
      type(c_ptr), value :: shift
      real(c_double), value :: delta, cellVol
    end function
    integer(c_int) function dftbp_b200_add_shift_atom(ctx, ham, over, shift, nSpin, isInputZero) bind(C)
      import :: c_ptr, c_int, c_double
      type(c_ptr), value :: ctx
      real(c_double), intent(inout) :: ham(*)
      real(c_double), intent(in) :: over(*), shift(*)
      integer(c_int), value :: nSpin, isInputZero
    end function
    integer(c_int) function dftbp_b200_add_shift_block(ctx, ham, over, shift, nSpin, isInputZero) bind(C)
      import :: c_ptr, c_int, c_double
      type(c_ptr), value :: ctx
      real(c_double), intent(inout) :: ham(*)
      real(c_double), intent(in) :: over(*), shift(*)
      integer(c_int), value :: nSpin, isInputZero
    end function
    integer(c_int) function dftbp_b200_derivative_shift(ctx, deriv, DM, nSpin, EDM, shift, diffType,&
        & delta) bind(C)
      import :: c_ptr, c_int, c_double
      type(c_ptr), value :: ctx
      real(c_double), intent(out) :: deriv(*)
      real(c_double), intent(in) :: DM(*), EDM(*)
      integer(c_int), value :: nSpin, diffType
      type(c_ptr), value :: shift
      real(c_double), value :: delta
    end function
    integer(c_int) function dftbp_b200_mulliken_atom(ctx, qq, over, rho) bind(C)
      import :: c_ptr, c_int, c_double
      type(c_ptr), value :: ctx
      real(c_double), intent(inout) :: qq(*)
      real(c_double), intent(in) :: over(*), rho(*)
    end function
    integer(c_int) function dftbp_b200_onsite_population(ctx, rho, qq) bind(C)
      import :: c_ptr, c_int, c_double
      type(c_ptr), value :: ctx
      real(c_double), intent(in) :: rho(*)
      real(c_double), intent(out) :: qq(*)
    end function
    integer(c_int) function dftbp_b200_fold_coords_to_cell(ctx, nAtom, coord, latVec) bind(C)
      import :: c_ptr, c_int, c_double
      type(c_ptr), value :: ctx
      integer(c_int), value :: nAtom
      real(c_double), intent(inout) :: coord(*)
      real(c_double), intent(in) :: latVec(3, 3)
    end function
    integer(c_int) function dftbp_b200_update_neighbour_list(ctx, nAtom, coord0, species0, nCellVec, cellVec,&
        & rCellVec, cutoff, skCutoff, symmetric, nAllAtom, mNeigh, sparseSize) bind(C)
      import :: c_ptr, c_int, c_double
      type(c_ptr), value :: ctx
      integer(c_int), value :: nAtom, nCellVec
      real(c_double), intent(in) :: coord0(*), cellVec(*), rCellVec(*)
      integer(c_int), intent(in) :: species0(*)
      real(c_double), value :: cutoff, skCutoff
      integer(c_int), value :: symmetric
      integer(c_int), intent(out) :: nAllAtom, mNeigh, sparseSize
    end function
    integer(c_int) function dftbp_b200_get_neighbour_list(ctx, coords, species, img2CentCell, iCellVec,&
        & nNeighbour, iNeighbour, neighDist2, nNeighbourSK, iPair, iAtomStart) bind(C)
      import :: c_ptr, c_int, c_double
      type(c_ptr), value :: ctx
      real(c_double), intent(out) :: coords(*), neighDist2(*)
      integer(c_int), intent(out) :: species(*), img2CentCell(*), iCellVec(*), nNeighbour(*), iNeighbour(*)
      integer(c_int), intent(out) :: nNeighbourSK(*), iPair(*)
      type(c_ptr), value :: iAtomStart
    end function
    integer(c_int) function dftbp_b200_nr_of_neighbours(ctx, cutoff, nNeighbour) bind(C)
      import :: c_ptr, c_int, c_double
      type(c_ptr), value :: ctx
      real(c_double), value :: cutoff
      integer(c_int), intent(out) :: nNeighbour(*)
    end function
    integer(c_int) function dftbp_b200_destroy(ctx) bind(C)
      import :: c_ptr, c_int
      type(c_ptr), value :: ctx
    end function
    integer(c_int) function dftbp_b200_update_coords(ctx, coords) bind(C)
      import :: c_ptr, c_int, c_double
      type(c_ptr), value :: ctx
      real(c_double), intent(in) :: coords(*)
    end function
    integer(c_int) function dftbp_b200_density_matrix_cplx(ctx, dm, eigvecs, n, ld, filling, weights)&
        & bind(C)
      import :: c_ptr, c_int, c_double, c_double_complex
      type(c_ptr), value :: ctx
      complex(c_double_complex), intent(out) :: dm(*)
      complex(c_double_complex), intent(in) :: eigvecs(*)
      real(c_double), intent(in) :: filling(*)
      integer(c_int), value :: n, ld
      type(c_ptr), value :: weights
    end function
    integer(c_int) function dftbp_b200_mulliken_block(ctx, qq, over, rho) bind(C)
      import :: c_ptr, c_int, c_double
      type(c_ptr), value :: ctx
      real(c_double), intent(inout) :: qq(*)
      real(c_double), intent(in) :: over(*), rho(*)
    end function
    integer(c_int) function dftbp_b200_hegv_real(ctx, H, S, eigen, n, ld, jobz) bind(C)
      import :: c_ptr, c_int, c_double, c_char
      type(c_ptr), value :: ctx
      real(c_double), intent(inout) :: H(*), S(*)
      real(c_double), intent(out) :: eigen(*)
      integer(c_int), value :: n, ld
      character(kind=c_char), value :: jobz
    end function
    integer(c_int) function dftbp_b200_hegv_cplx(ctx, H, S, eigen, n, ld, jobz) bind(C)
      import :: c_ptr, c_int, c_double, c_double_complex, c_char
      type(c_ptr), value :: ctx
      complex(c_double_complex), intent(inout) :: H(*), S(*)
      real(c_double), intent(out) :: eigen(*)
      integer(c_int), value :: n, ld
      character(kind=c_char), value :: jobz
    end function
    integer(c_int) function dftbp_b200_efilling(ctx, Ebs, Ef, TS, E0, filling, eigenvals, nLevels, nKpt, nSpin,&
        & nElectrons, kT, kWeight, distrib) bind(C)
      import :: c_ptr, c_int, c_double
      type(c_ptr), value :: ctx
      real(c_double), intent(out) :: Ebs(*), Ef, TS(*), E0(*), filling(*)
      real(c_double), intent(in) :: eigenvals(*), kWeight(*)
      integer(c_int), value :: nLevels, nKpt, nSpin, distrib
      real(c_double), value :: nElectrons, kT
    end function
    integer(c_int) function dftbp_b200_invr_cluster(ctx, nAtom, coord, invRMat, ld) bind(C)
      import :: c_ptr, c_int, c_double
      type(c_ptr), value :: ctx
      integer(c_int), value :: nAtom, ld
      real(c_double), intent(in) :: coord(*)
      real(c_double), intent(out) :: invRMat(*)
    end function
    integer(c_int) function dftbp_b200_invr_periodic(ctx, nAtom, coord, nRLat, rLat, nG, gLat, alpha, volume,&
        & rCutoff, invRMat, ld) bind(C)
      import :: c_ptr, c_int, c_double
      type(c_ptr), value :: ctx
      integer(c_int), value :: nAtom, nRLat, nG, ld
      real(c_double), intent(in) :: coord(*), rLat(*), gLat(*)
      real(c_double), value :: alpha, volume, rCutoff
      real(c_double), intent(out) :: invRMat(*)
    end function
    integer(c_int) function dftbp_b200_coulomb_shift(ctx, nAtom, invRMat, ld, deltaQAtom, shiftPerAtom) bind(C)
      import :: c_ptr, c_int, c_double
      type(c_ptr), value :: ctx
      integer(c_int), value :: nAtom, ld
      real(c_double), intent(in) :: invRMat(*), deltaQAtom(*)
      real(c_double), intent(out) :: shiftPerAtom(*)
    end function
    integer(c_int) function dftbp_b200_short_gamma_update(ctx, nSpecies, mHubbU, nHubbU, uniqHubbU, shortCutoffs)&
        & bind(C)
      import :: c_ptr, c_int, c_double
      type(c_ptr), value :: ctx
      integer(c_int), value :: nSpecies, mHubbU
      integer(c_int), intent(in) :: nHubbU(*)
      real(c_double), intent(in) :: uniqHubbU(*), shortCutoffs(*)
    end function
    integer(c_int) function dftbp_b200_short_gamma_shifts(ctx, iHubbU, deltaQShell, shiftShell) bind(C)
      import :: c_ptr, c_int, c_double
      type(c_ptr), value :: ctx
      integer(c_int), intent(in) :: iHubbU(*)
      real(c_double), intent(in) :: deltaQShell(*)
      real(c_double), intent(out) :: shiftShell(*)
    end function
    integer(c_int) function dftbp_b200_mixer_init(ctx, maxSccIter, mixParam, omega0, minWeight, maxWeight,&
        & weightFac) bind(C)
      import :: c_ptr, c_int, c_double
      type(c_ptr), value :: ctx
      integer(c_int), value :: maxSccIter
      real(c_double), value :: mixParam, omega0, minWeight, maxWeight, weightFac
    end function
    integer(c_int) function dftbp_b200_mixer_reset(ctx, nElem) bind(C)
      import :: c_ptr, c_int
      type(c_ptr), value :: ctx
      integer(c_int), value :: nElem
    end function
    integer(c_int) function dftbp_b200_mixer_mix(ctx, qInpResult, qDiff, nElem) bind(C)
      import :: c_ptr, c_int, c_double
      type(c_ptr), value :: ctx
      real(c_double), intent(inout) :: qInpResult(*)
      real(c_double), intent(in) :: qDiff(*)
      integer(c_int), value :: nElem
    end function
    integer(c_int) function dftbp_b200_adopt_neighbour_list(ctx) bind(C)
      import :: c_ptr, c_int
      type(c_ptr), value :: ctx
    end function
  end interface

contains

  !> Non-zero status -> the reference's abort path (io/message.F90:74-84).  Routines that carry a
  !> TStatus (getDensityMatrix) would use @:RAISE_ERROR(errStatus, rc, msg) instead.
  subroutine check(ctx, rc)
    use dftbp_io_message, only : error
    type(TB200Ctx), intent(in) :: ctx
    integer(c_int), intent(in) :: rc
    character(len=512) :: msg
    if (rc == 0) return
    call lastError(ctx, msg)
    call error("libdftbp_b200: " // trim(msg))
  end subroutine check

  !> Message of the last failed call
  subroutine lastError(ctx, msg)
    type(TB200Ctx), intent(in) :: ctx
    character(len=*), intent(out) :: msg
    character(kind=c_char), pointer :: cmsg(:)
    integer :: ii
    call c_f_pointer(dftbp_b200_last_error(ctx%h), cmsg, [len(msg)])
    msg = ""
    do ii = 1, len(msg)
      if (cmsg(ii) == c_null_char) exit
      msg(ii:ii) = cmsg(ii)
    end do
  end subroutine lastError

  !> Non-zero status -> TStatus, for routines that carry one (what @:RAISE_ERROR expands to,
  !> include/error.fypp:87-90, without the return)
  subroutine checkStatus(ctx, rc, errStatus, line)
    use dftbp_common_status, only : TStatus
    type(TB200Ctx), intent(in) :: ctx
    integer(c_int), intent(in) :: rc
    type(TStatus), intent(inout) :: errStatus
    integer, intent(in) :: line
    character(len=512) :: msg
    if (rc == 0) return
    call lastError(ctx, msg)
    ! the reference raises code -1 with "Atoms i and j too close together" (periodic.F90:645-648)
    call errStatus%setError(-1, trim(msg), "dftbp_b200_shim.F90", line)
  end subroutine checkStatus

  !> Once per run, after readSKFiles (dftbplus/parser.F90:3668): flattens TOrbitals + TSlakoCont.
  !> skHam/skOver(iSp1,iSp2) hold the arrays that were passed to SlakoEqGrid_init.
  subroutine b200_init(ctx, device, rank, nRanks, orb, nGrid, dist, skHam, skOver, skIntMethod)
    use dftbp_type_orbitals, only : TOrbitals
    type(TB200Ctx), intent(out) :: ctx
    integer, intent(in) :: device, rank, nRanks
    type(TOrbitals), intent(in) :: orb
    integer, intent(in) :: nGrid(:,:), skIntMethod
    real(dp), intent(in) :: dist(:,:)
    real(dp), intent(in) :: skHam(:,:,:,:), skOver(:,:,:,:)   ! (nGridMax, nIntegMax, iSp1, iSp2)
    integer :: iSp1, iSp2, nSp, nInt
    nSp = size(orb%nShell)
    call check(ctx, dftbp_b200_create(ctx%h, int(device, c_int), int(rank, c_int), int(nRanks, c_int)))
    call check(ctx, dftbp_b200_set_orbitals(ctx%h, int(nSp, c_int), int(orb%mShell, c_int), orb%nShell,&
        & orb%angShell))
    do iSp1 = 1, nSp
      do iSp2 = 1, nSp
        nInt = nIntegrals(orb, iSp1, iSp2)
        call check(ctx, dftbp_b200_set_sk_table(ctx%h, 0_c_int, iSp1, iSp2, nGrid(iSp1,iSp2), nInt,&
            & dist(iSp1,iSp2), skHam(1:nGrid(iSp1,iSp2), 1:nInt, iSp1, iSp2), skIntMethod))
        call check(ctx, dftbp_b200_set_sk_table(ctx%h, 1_c_int, iSp1, iSp2, nGrid(iSp1,iSp2), nInt,&
            & dist(iSp1,iSp2), skOver(1:nGrid(iSp1,iSp2), 1:nInt, iSp1, iSp2), skIntMethod))
      end do
    end do
  contains
    integer function nIntegrals(orb, iSp1, iSp2)
      type(TOrbitals), intent(in) :: orb
      integer, intent(in) :: iSp1, iSp2
      integer :: iSh1, iSh2
      nIntegrals = 0
      do iSh1 = 1, orb%nShell(iSp1)
        do iSh2 = 1, orb%nShell(iSp2)
          nIntegrals = nIntegrals + min(orb%angShell(iSh1, iSp1), orb%angShell(iSh2, iSp2)) + 1
        end do
      end do
    end function nIntegrals
  end subroutine b200_init

  !> Once per geometry, at the end of handleCoordinateChange (dftbplus/main.F90:2413-2577).
  subroutine b200_setGeometry(ctx, coords, species, img2CentCell, iCellVec, cellVec, nNeighbourSK,&
      & iNeighbour, iSparseStart, iAtomStart, sparseSize)
    type(TB200Ctx), intent(in) :: ctx
    real(dp), intent(in) :: coords(:,:), cellVec(:,:)
    integer, intent(in) :: species(:), img2CentCell(:), iCellVec(:), nNeighbourSK(:)
    integer, intent(in) :: iNeighbour(0:,:), iSparseStart(0:,:), iAtomStart(:), sparseSize
    integer :: nAtom, nAllAtom, mNeigh
    nAtom = size(nNeighbourSK)
    nAllAtom = size(coords, dim=2)
    mNeigh = ubound(iNeighbour, dim=1)
    ! iSparseStart may be allocated wider than iNeighbour; pass matching (0:mNeigh, nAtom) sections
    call check(ctx, dftbp_b200_set_geometry(ctx%h, nAtom, nAllAtom, mNeigh, coords, species,&
        & img2CentCell, iCellVec, size(cellVec, dim=2), cellVec, nNeighbourSK, iNeighbour,&
        & iSparseStart(0:mNeigh, :), iAtomStart, sparseSize))
  end subroutine b200_setGeometry

  !> getNrOfNeighboursForAll, dftb/periodic.F90:996 (another cutoff on the device list)
  subroutine b200_getNrOfNeighboursForAll(ctx, nNeighbour, cutoff)
    type(TB200Ctx), intent(in) :: ctx
    integer, intent(out) :: nNeighbour(:)
    real(dp), intent(in) :: cutoff
    call check(ctx, dftbp_b200_nr_of_neighbours(ctx%h, cutoff, nNeighbour))
  end subroutine b200_getNrOfNeighboursForAll

  !> boundaryCond%foldCoordsToCell, pbc3d without contacts (dftb/boundarycond.F90:369-380)
  subroutine b200_foldCoordsToCell(ctx, coord, latVec)
    type(TB200Ctx), intent(in) :: ctx
    real(dp), intent(inout) :: coord(:,:)
    real(dp), intent(in) :: latVec(:,:)
    call check(ctx, dftbp_b200_fold_coords_to_cell(ctx%h, size(coord, dim=2), coord, latVec))
  end subroutine b200_foldCoordsToCell

  !> updateNeighbourListAndSpecies (dftb/periodic.F90:380) + getNrOfNeighboursForAll (:996) +
  !> getSparseDescriptor (dftb/sparse2dense.F90:3617) in one device pass; replaces the three calls of
  !> handleCoordinateChange (dftbplus/main.F90:2562-2578) and the later b200_setGeometry.  The
  !> arrays come back with the reference's shapes (rows 0:maxval(nNeighbour)) for the host-side
  !> consumers (repulsive, SCC, dispersion).
  subroutine b200_updateNeighbourListAndSpecies(ctx, coord, species, img2CentCell, iCellVec, neigh,&
      & nAllAtom, coord0, species0, cutoff, skCutoff, cellVec, rCellVec, nNeighbourSK, iSparseStart,&
      & sparseSize, errStatus)
    use dftbp_common_status, only : TStatus
    use dftbp_dftb_periodic, only : TNeighbourList, allocateNeighbourArrays
    type(TB200Ctx), intent(in) :: ctx
    real(dp), allocatable, intent(inout) :: coord(:,:)
    integer, allocatable, intent(inout) :: species(:), img2CentCell(:), iCellVec(:), iSparseStart(:,:)
    type(TNeighbourList), intent(inout) :: neigh
    integer, intent(out) :: nAllAtom, sparseSize
    real(dp), intent(in) :: coord0(:,:), cutoff, skCutoff, cellVec(:,:), rCellVec(:,:)
    integer, intent(in) :: species0(:)
    integer, intent(out) :: nNeighbourSK(:)
    type(TStatus), intent(out) :: errStatus
    integer :: nAtom, mNeigh
    if (neigh%setExternally) return
    nAtom = size(species0)
    call checkStatus(ctx, dftbp_b200_update_neighbour_list(ctx%h, nAtom, coord0, species0,&
        & size(rCellVec, dim=2), cellVec, rCellVec, cutoff, skCutoff, 0_c_int, nAllAtom, mNeigh, sparseSize),&
        & errStatus,&
        & __LINE__)
    if (errStatus%hasError()) return   ! @:PROPAGATE_ERROR(errStatus) at the call site, main.F90:2572
    if (allocated(coord)) deallocate(coord, species, img2CentCell, iCellVec)
    allocate(coord(3, nAllAtom), species(nAllAtom), img2CentCell(nAllAtom), iCellVec(nAllAtom))
    ! iNeighbour / neighDist2 are pointers into memory the list owns (periodic.F90:881-922)
    call allocateNeighbourArrays(neigh, mNeigh, nAtom)
    if (allocated(iSparseStart)) deallocate(iSparseStart)
    allocate(iSparseStart(0:mNeigh, nAtom))
    neigh%cutoff = cutoff
    call check(ctx, dftbp_b200_get_neighbour_list(ctx%h, coord, species, img2CentCell, iCellVec,&
        & neigh%nNeighbour, neigh%iNeighbour, neigh%neighDist2, nNeighbourSK, iSparseStart, c_null_ptr))
    call check(ctx, dftbp_b200_adopt_neighbour_list(ctx%h))
  end subroutine b200_updateNeighbourListAndSpecies

  !> buildH0, dftb/nonscc.F90:66 (env, skHamCont, coords, ... are held by the context)
  subroutine b200_buildH0(ctx, ham, selfegy)
    type(TB200Ctx), intent(in) :: ctx
    real(dp), intent(out) :: ham(:)
    real(dp), intent(in) :: selfegy(:,:)
    call check(ctx, dftbp_b200_build_h0(ctx%h, selfegy, ham))
  end subroutine b200_buildH0

  !> buildS, dftb/nonscc.F90:128
  subroutine b200_buildS(ctx, over)
    type(TB200Ctx), intent(in) :: ctx
    real(dp), intent(out) :: over(:)
    call check(ctx, dftbp_b200_build_s(ctx%h, over))
  end subroutine b200_buildS

  !> unpackHS_real, dftb/sparse2dense.F90:266
  subroutine b200_unpackHS_real(ctx, square, orig)
    type(TB200Ctx), intent(in) :: ctx
    real(dp), intent(out) :: square(:,:)
    real(dp), intent(in) :: orig(:)
    call check(ctx, dftbp_b200_unpack_hs_real(ctx%h, orig, square, size(square, dim=1)))
  end subroutine b200_unpackHS_real

  !> unpackHS_cmplx_kpts, dftb/sparse2dense.F90:185
  subroutine b200_unpackHS_cmplx(ctx, square, orig, kPoint)
    type(TB200Ctx), intent(in) :: ctx
    complex(dp), intent(out) :: square(:,:)
    real(dp), intent(in) :: orig(:), kPoint(:)
    call check(ctx, dftbp_b200_unpack_hs_cplx(ctx%h, orig, kPoint, square, size(square, dim=1)))
  end subroutine b200_unpackHS_cmplx

  !> packHS_real, dftb/sparse2dense.F90:1024
  subroutine b200_packHS_real(ctx, primitive, square)
    type(TB200Ctx), intent(in) :: ctx
    real(dp), intent(inout) :: primitive(:)
    real(dp), intent(in) :: square(:,:)
    call check(ctx, dftbp_b200_pack_hs_real(ctx%h, primitive, square, size(square, dim=1)))
  end subroutine b200_packHS_real

  !> packHS_cmplx_kpts, dftb/sparse2dense.F90:820
  subroutine b200_packHS_cmplx(ctx, primitive, square, kPoint, kWeight)
    type(TB200Ctx), intent(in) :: ctx
    real(dp), intent(inout) :: primitive(:)
    complex(dp), intent(in) :: square(:,:)
    real(dp), intent(in) :: kPoint(:), kWeight
    call check(ctx, dftbp_b200_pack_hs_cplx(ctx%h, primitive, square, size(square, dim=1), kPoint, kWeight))
  end subroutine b200_packHS_cmplx

  !> fullDensityMatrix_real, dftb/densitymatrix.F90:530 (eigenvecs is left untouched)
  subroutine b200_getDensityMatrix_real(ctx, dm, eigenvecs, filling)
    type(TB200Ctx), intent(in) :: ctx
    real(dp), intent(out) :: dm(:,:)
    real(dp), intent(in) :: eigenvecs(:,:), filling(:)
    call check(ctx, dftbp_b200_density_matrix_real(ctx%h, dm, eigenvecs, size(dm, dim=2),&
        & size(dm, dim=1), filling, c_null_ptr))
  end subroutine b200_getDensityMatrix_real

  !> fullDensityMatrix_cmplx, dftb/densitymatrix.F90:597 (caller: getDensityFromCplxEigvecs,
  !> dftbplus/main.F90:4289 -- one call per local (k-point, spin) of TParallelKS)
  subroutine b200_getDensityMatrix_cmplx(ctx, dm, eigenvecs, filling)
    type(TB200Ctx), intent(in) :: ctx
    complex(dp), intent(out) :: dm(:,:)
    complex(dp), intent(in) :: eigenvecs(:,:)
    real(dp), intent(in) :: filling(:)
    call check(ctx, dftbp_b200_density_matrix_cplx(ctx%h, dm, eigenvecs, size(dm, dim=2),&
        & size(dm, dim=1), filling, c_null_ptr))
  end subroutine b200_getDensityMatrix_cmplx

  !> fullEnergyDensityMatrix_real, dftb/densitymatrix.F90:664 (caller: getEDensityMtxFromRealEigvecs,
  !> dftbplus/main.F90:6572): dm = C diag(filling * eigen) C^T; filling decides the number of levels
  subroutine b200_getEDensityMatrix_real(ctx, dm, eigenvecs, filling, eigen)
    type(TB200Ctx), intent(in) :: ctx
    real(dp), intent(out) :: dm(:,:)
    real(dp), intent(in) :: eigenvecs(:,:), filling(:), eigen(:)
    real(dp), allocatable, target :: weights(:)
    weights = filling * eigen
    call check(ctx, dftbp_b200_density_matrix_real(ctx%h, dm, eigenvecs, size(dm, dim=2),&
        & size(dm, dim=1), filling, c_loc(weights)))
  end subroutine b200_getEDensityMatrix_real

  !> fullEnergyDensityMatrix_cmplx, dftb/densitymatrix.F90:739 (caller: getEDensityMtxFromComplexEigvecs,
  !> dftbplus/main.F90:6808)
  subroutine b200_getEDensityMatrix_cmplx(ctx, dm, eigenvecs, filling, eigen)
    type(TB200Ctx), intent(in) :: ctx
    complex(dp), intent(out) :: dm(:,:)
    complex(dp), intent(in) :: eigenvecs(:,:)
    real(dp), intent(in) :: filling(:), eigen(:)
    real(dp), allocatable, target :: weights(:)
    weights = filling * eigen
    call check(ctx, dftbp_b200_density_matrix_cplx(ctx%h, dm, eigenvecs, size(dm, dim=2),&
        & size(dm, dim=1), filling, c_loc(weights)))
  end subroutine b200_getEDensityMatrix_cmplx

  !> diagDenseMtx, dftbplus/eigenvects.F90:85-100 (hegvd branch, real): on exit HSqr holds the
  !> eigenvectors, SSqr its Cholesky factor
  subroutine b200_hegv_real(ctx, HSqr, SSqr, eigen, jobz)
    type(TB200Ctx), intent(in) :: ctx
    real(dp), intent(inout) :: HSqr(:,:), SSqr(:,:)
    real(dp), intent(out) :: eigen(:)
    character, intent(in) :: jobz
    call check(ctx, dftbp_b200_hegv_real(ctx%h, HSqr, SSqr, eigen, size(HSqr, dim=2), size(HSqr, dim=1),&
        & jobz))
  end subroutine b200_hegv_real

  !> diagDenseMtx, complex Hermitian case (k-points)
  subroutine b200_hegv_cmplx(ctx, HSqr, SSqr, eigen, jobz)
    type(TB200Ctx), intent(in) :: ctx
    complex(dp), intent(inout) :: HSqr(:,:), SSqr(:,:)
    real(dp), intent(out) :: eigen(:)
    character, intent(in) :: jobz
    call check(ctx, dftbp_b200_hegv_cplx(ctx%h, HSqr, SSqr, eigen, size(HSqr, dim=2), size(HSqr, dim=1),&
        & jobz))
  end subroutine b200_hegv_cmplx

  !> Efilling, dftb/etemp.F90:59 (same dummy arguments; Fermi distribution only -- other values of distrib
  !> stay with the Fortran routine)
  subroutine b200_Efilling(ctx, Ebs, Ef, TS, E0, filling, eigenvals, nElectrons, kT, kWeight, distrib)
    type(TB200Ctx), intent(in) :: ctx
    real(dp), intent(out) :: Ebs(:), Ef, TS(:), E0(:), filling(:,:,:)
    real(dp), intent(in) :: eigenvals(:,:,:), nElectrons, kT, kWeight(:)
    integer, intent(in) :: distrib
    call check(ctx, dftbp_b200_efilling(ctx%h, Ebs, Ef, TS, E0, filling, eigenvals, size(eigenvals, dim=1),&
        & size(eigenvals, dim=2), size(eigenvals, dim=3), nElectrons, kT, kWeight, int(distrib, c_int)))
  end subroutine b200_Efilling

  !> invRCluster, dftb/coulomb.F90:623 (serial branch :704): coord(3, nAtom), invRMat(nAtom, nAtom)
  subroutine b200_invRCluster(ctx, nAtom, coord, invRMat)
    type(TB200Ctx), intent(in) :: ctx
    integer, intent(in) :: nAtom
    real(dp), intent(in) :: coord(:,:)
    real(dp), intent(out) :: invRMat(:,:)
    call check(ctx, dftbp_b200_invr_cluster(ctx%h, int(nAtom, c_int), coord, invRMat, size(invRMat, dim=1)))
  end subroutine b200_invRCluster

  !> invRPeriodic, dftb/coulomb.F90:815 (serial branch :968).  The dynamic neighbour list argument of the
  !> reference is replaced by what defines it: the lattice translations this%rLatPoints_ (origin first) and
  !> the real-space cutoff maxREwald, both from TCoulomb%updateLatVecs (:330-372)
  subroutine b200_invRPeriodic(ctx, nAtom, coord, rLatPoints, maxREwald, recPoint, alpha, volume, invRMat)
    type(TB200Ctx), intent(in) :: ctx
    integer, intent(in) :: nAtom
    real(dp), intent(in) :: coord(:,:), rLatPoints(:,:), maxREwald, recPoint(:,:), alpha, volume
    real(dp), intent(out) :: invRMat(:,:)
    call check(ctx, dftbp_b200_invr_periodic(ctx%h, int(nAtom, c_int), coord, size(rLatPoints, dim=2), rLatPoints,&
        & size(recPoint, dim=2), recPoint, alpha, volume, maxREwald, invRMat, size(invRMat, dim=1)))
  end subroutine b200_invRPeriodic

  !> TCoulomb%updateShifts, dftb/coulomb.F90:537: shiftPerAtom = hemv(invRMat, deltaQAtom, 'L')
  subroutine b200_coulombShifts(ctx, invRMat, deltaQAtom, shiftPerAtom)
    type(TB200Ctx), intent(in) :: ctx
    real(dp), intent(in) :: invRMat(:,:), deltaQAtom(:)
    real(dp), intent(out) :: shiftPerAtom(:)
    call check(ctx, dftbp_b200_coulomb_shift(ctx%h, size(deltaQAtom), invRMat, size(invRMat, dim=1), deltaQAtom,&
        & shiftPerAtom))
  end subroutine b200_coulombShifts

  !> TShortGamma%updateCoords, dftb/shortgamma.F90:177 (undamped, no H5), on the device neighbour list of the
  !> last b200_updateNeighbourListAndSpecies: nHubbU / uniqHubbU = this%hubbU_, shortCutoffs = this%shortCutoffs_
  subroutine b200_shortGammaUpdateCoords(ctx, nHubbU, uniqHubbU, shortCutoffs)
    type(TB200Ctx), intent(in) :: ctx
    integer, intent(in) :: nHubbU(:)
    real(dp), intent(in) :: uniqHubbU(:,:), shortCutoffs(:,:,:,:)
    call check(ctx, dftbp_b200_short_gamma_update(ctx%h, size(nHubbU), size(shortCutoffs, dim=1),&
        & int(nHubbU, c_int), uniqHubbU(1:size(shortCutoffs, dim=1), :), shortCutoffs))
  end subroutine b200_shortGammaUpdateCoords

  !> TShortGamma%updateShifts, dftb/shortgamma.F90:220: shiftShell(mShell, nAtom) = buildShifts_(deltaQShell)
  subroutine b200_shortGammaShifts(ctx, iHubbU, deltaQShell, shiftShell)
    type(TB200Ctx), intent(in) :: ctx
    integer, intent(in) :: iHubbU(:,:)
    real(dp), intent(in) :: deltaQShell(:,:)
    real(dp), intent(out) :: shiftShell(:,:)
    call check(ctx, dftbp_b200_short_gamma_shifts(ctx%h, int(iHubbU, c_int), deltaQShell, shiftShell))
  end subroutine b200_shortGammaShifts

  !> TBroydenMixerReal_init, mixer/broydenmixer.F90:110 (same dummy arguments after the instance)
  subroutine b200_BroydenMixer_init(ctx, mIter, mixParam, omega0, minWeight, maxWeight, weightFac)
    type(TB200Ctx), intent(in) :: ctx
    integer, intent(in) :: mIter
    real(dp), intent(in) :: mixParam, omega0, minWeight, maxWeight, weightFac
    call check(ctx, dftbp_b200_mixer_init(ctx%h, int(mIter, c_int), mixParam, omega0, minWeight, maxWeight, weightFac))
  end subroutine b200_BroydenMixer_init

  !> TBroydenMixerReal_reset, mixer/broydenmixer.F90:143
  subroutine b200_BroydenMixer_reset(ctx, nElem)
    type(TB200Ctx), intent(in) :: ctx
    integer, intent(in) :: nElem
    call check(ctx, dftbp_b200_mixer_reset(ctx%h, int(nElem, c_int)))
  end subroutine b200_BroydenMixer_reset

  !> TBroydenMixerReal_mix, mixer/broydenmixer.F90:180
  subroutine b200_BroydenMixer_mix(ctx, qInpResult, qDiff)
    type(TB200Ctx), intent(in) :: ctx
    real(dp), intent(inout) :: qInpResult(:)
    real(dp), intent(in) :: qDiff(:)
    call check(ctx, dftbp_b200_mixer_mix(ctx%h, qInpResult, qDiff, size(qInpResult)))
  end subroutine b200_BroydenMixer_mix

  !> mullikenPerBlock, dftb/populations.F90:306 -- qq(mOrb, mOrb, nAtom) accumulates
  subroutine b200_mullikenPerBlock(ctx, qq, over, rho)
    type(TB200Ctx), intent(in) :: ctx
    real(dp), intent(inout) :: qq(:,:,:)
    real(dp), intent(in) :: over(:), rho(:)
    call check(ctx, dftbp_b200_mulliken_block(ctx%h, qq, over, rho))
  end subroutine b200_mullikenPerBlock

  !> New coordinates on an unchanged neighbour list (geometry steps between list rebuilds):
  !> coords(3, nAllAtom) as updateNeighbourListAndSpecies left them
  subroutine b200_updateCoords(ctx, coords)
    type(TB200Ctx), intent(in) :: ctx
    real(dp), intent(in) :: coords(:,:)
    call check(ctx, dftbp_b200_update_coords(ctx%h, coords))
  end subroutine b200_updateCoords

  !> End of run (the counterpart of b200_init)
  subroutine b200_destroy(ctx)
    type(TB200Ctx), intent(inout) :: ctx
    integer(c_int) :: rc
    if (c_associated(ctx%h)) rc = dftbp_b200_destroy(ctx%h)
    ctx%h = c_null_ptr
  end subroutine b200_destroy

  !> mullikenPerOrbital, dftb/populations.F90:120
  subroutine b200_mullikenPerOrbital(ctx, qq, over, rho)
    type(TB200Ctx), intent(in) :: ctx
    real(dp), intent(inout) :: qq(:,:)
    real(dp), intent(in) :: over(:), rho(:)
    call check(ctx, dftbp_b200_mulliken_orbital(ctx%h, qq, over, rho))
  end subroutine b200_mullikenPerOrbital

  !> mullikenPerAtom, dftb/populations.F90:68 (the rank-1 overload of the generic mulliken)
  subroutine b200_mullikenPerAtom(ctx, qq, over, rho)
    type(TB200Ctx), intent(in) :: ctx
    real(dp), intent(inout) :: qq(:)
    real(dp), intent(in) :: over(:), rho(:)
    call check(ctx, dftbp_b200_mulliken_atom(ctx%h, qq, over, rho))
  end subroutine b200_mullikenPerAtom

  !> getOnsitePopulation, dftb/populations.F90:1195
  subroutine b200_getOnsitePopulation(ctx, rho, qq)
    type(TB200Ctx), intent(in) :: ctx
    real(dp), intent(in) :: rho(:)
    real(dp), intent(out) :: qq(:)
    call check(ctx, dftbp_b200_onsite_population(ctx%h, rho, qq))
  end subroutine b200_getOnsitePopulation

  !> getBlockStress, dftb/stress.F90:188 (getNonSCCStress, :70, when shift is absent)
  subroutine b200_getBlockStress(ctx, st, DM, EDM, shift, diffType, deltaXDiff, cellVol)
    type(TB200Ctx), intent(in) :: ctx
    real(dp), intent(out) :: st(:,:)
    real(dp), intent(in), target :: DM(:,:), EDM(:)
    real(dp), intent(in), target, optional :: shift(:,:,:,:)
    integer, intent(in) :: diffType
    real(dp), intent(in) :: deltaXDiff, cellVol
    type(c_ptr) :: pShift
    pShift = c_null_ptr
    if (present(shift)) pShift = c_loc(shift)
    call check(ctx, dftbp_b200_stress(ctx%h, st, DM, size(DM, dim=2), EDM, pShift, diffType, deltaXDiff,&
        & cellVol))
  end subroutine b200_getBlockStress

  !> addShift_atom, dftb/shift.F90:41 -- shift(nAtom, nSpin)
  subroutine b200_addShift_atom(ctx, ham, over, shift, isInputZero)
    type(TB200Ctx), intent(in) :: ctx
    real(dp), intent(inout) :: ham(:,:)
    real(dp), intent(in) :: over(:), shift(:,:)
    logical, intent(in) :: isInputZero
    call check(ctx, dftbp_b200_add_shift_atom(ctx%h, ham, over, shift, size(ham, dim=2),&
        & merge(1_c_int, 0_c_int, isInputZero)))
  end subroutine b200_addShift_atom

  !> addShift_block, dftb/shift.F90:132 -- shift(mOrb, mOrb, nAtom, nSpin)
  subroutine b200_addShift_block(ctx, ham, over, shift, isInputZero)
    type(TB200Ctx), intent(in) :: ctx
    real(dp), intent(inout) :: ham(:,:)
    real(dp), intent(in) :: over(:), shift(:,:,:,:)
    logical, intent(in) :: isInputZero
    call check(ctx, dftbp_b200_add_shift_block(ctx%h, ham, over, shift, size(ham, dim=2),&
        & merge(1_c_int, 0_c_int, isInputZero)))
  end subroutine b200_addShift_block

  !> derivative_shift, dftb/forces.F90:27-38 (block variant when shift is present)
  subroutine b200_derivative_shift(ctx, deriv, DM, EDM, shift, diffType, deltaXDiff)
    type(TB200Ctx), intent(in) :: ctx
    real(dp), intent(out) :: deriv(:,:)
    real(dp), intent(in), target :: DM(:,:)
    real(dp), intent(in) :: EDM(:)
    real(dp), intent(in), target, optional :: shift(:,:,:,:)
    integer, intent(in) :: diffType
    real(dp), intent(in) :: deltaXDiff
    type(c_ptr) :: pShift
    pShift = c_null_ptr
    if (present(shift)) pShift = c_loc(shift)
    call check(ctx, dftbp_b200_derivative_shift(ctx%h, deriv, DM, size(DM, dim=2), EDM, pShift,&
        & int(diffType, c_int), deltaXDiff))
  end subroutine b200_derivative_shift

end module dftbp_b200_shim

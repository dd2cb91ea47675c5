! cpb200_interfaces: iso_c_binding view of include/cpb200.h for the CPMD source tree.
!
! To be added next to the reference's other C-binding modules (cuuser_interfaces.mod.F90,
! cufft_interfaces.mod.F90).  NOT compile-tested in the authoring image (no Fortran compiler):
! the argument kinds and order below are mirrored one to one by cpmd_b200/lib.py (ctypes), which
! tests/test_abi.py checks against the header and the shared object, and the header itself is
! compiled as C99 in the same test.  INTEGRATION.md shows where the calls go.
!
! Conventions: assumed-shape dummies never cross the boundary - the callers pass C_LOC(a(1,1)) or an
! explicit-size view plus the leading dimension; LOGICALs stay on the Fortran side; every function
! returns 0 on success, otherwise CALL stopgm(procedureN, cpb_error_message(), __LINE__, __FILE__).
MODULE cpb200_interfaces
  USE, INTRINSIC :: iso_c_binding, ONLY: c_int, c_long, c_size_t, c_double, c_ptr, c_char, c_null_char, &
       c_f_pointer, c_associated
  IMPLICIT NONE
  PRIVATE

  ! flags (include/cpb200.h)
  INTEGER(c_int), PARAMETER, PUBLIC :: CPB_VPSI_OVERWRITE = 1, CPB_VPSI_TKSHAM = 2
  INTEGER(c_int), PARAMETER, PUBLIC :: CPB_RHO_CHECK_CHARGE = 1, CPB_RHO_ACCUMULATE = 2
  INTEGER(c_int), PARAMETER, PUBLIC :: CPB_C0_KEEP = 16, CPB_C0_REUSE = 32
  INTEGER(c_int), PARAMETER, PUBLIC :: CPB_PSI_KEEP = 64, CPB_PSI_REUSE = 128
  INTEGER, PARAMETER, PUBLIC :: CPB_PEER_HANDLE_BYTES = 64

  PUBLIC :: cpb_length_supported, cpb_plan_create, cpb_plan_destroy
  PUBLIC :: cpb_rhoofr, cpb_vpsi, cpb_rhoofr_lsd, cpb_vpsi_lsd, cpb_c0_invalidate
  PUBLIC :: cpb_rhoofr_kpt, cpb_vpsi_kpt, cpb_tauofr, cpb_vtaupsi, cpb_vofrho_local, cpb_hfx
  PUBLIC :: cpb_peer_create, cpb_peer_connect, cpb_peer_local_ptr, cpb_peer_allreduce_f64, &
       cpb_peer_bcast_f64, cpb_peer_check, cpb_peer_destroy, cpb_peer_redist_c2, cpb_peer_allgather_f64, &
       cpb_peer_allreduce_scalars, cpb_peer_set_timeout_ms
  PUBLIC :: cpb_error_message

  INTERFACE
     INTEGER(c_int) FUNCTION cpb_length_supported(n) BIND(c, name='cpb_length_supported')
       IMPORT :: c_int
       INTEGER(c_int), VALUE :: n
     END FUNCTION cpb_length_supported

     ! nr = spar%nr1s..nr3s, kr = fpar%kr1,kr2s,kr3s, inyh/hg from cppt, tpiba2/omega from parm.
     ! Wavefunction plan: ngw = ncpw%ngw.  Density plan (cpb_vofrho_local): ngw = ncpw%nhg.
     INTEGER(c_int) FUNCTION cpb_plan_create(plan, nr, kr, ngw, inyh, hg, tpiba2, omega, device, &
          max_batch_pairs) BIND(c, name='cpb_plan_create')
       IMPORT :: c_int, c_ptr, c_double
       TYPE(c_ptr), INTENT(out) :: plan
       INTEGER(c_int), INTENT(in) :: nr(3), kr(3)
       INTEGER(c_int), VALUE :: ngw
       INTEGER(c_int), INTENT(in) :: inyh(3,*)
       REAL(c_double), INTENT(in) :: hg(*)
       REAL(c_double), VALUE :: tpiba2, omega
       INTEGER(c_int), VALUE :: device, max_batch_pairs
     END FUNCTION cpb_plan_create

     INTEGER(c_int) FUNCTION cpb_plan_destroy(plan) BIND(c, name='cpb_plan_destroy')
       IMPORT :: c_int, c_ptr
       TYPE(c_ptr), VALUE :: plan
     END FUNCTION cpb_plan_destroy

     ! SUBROUTINE rhoofr(c0,rhoe,psi,nstate)                         rhoofr_utils.mod.F90:122
     INTEGER(c_int) FUNCTION cpb_rhoofr(plan, c0, ld_c0, nstate, f, ngroups, my_group, rhoe, ekin, &
          rsum_g, rsum_r, flags) BIND(c, name='cpb_rhoofr')
       IMPORT :: c_int, c_long, c_ptr, c_double
       TYPE(c_ptr), VALUE :: plan, c0                    ! c0 = C_LOC(c0(1,1)), COMPLEX(real_8)
       INTEGER(c_long), VALUE :: ld_c0                   ! SIZE(c0,1)
       INTEGER(c_int), VALUE :: nstate
       REAL(c_double), INTENT(in) :: f(*)                ! crge%f(:,1)
       INTEGER(c_int), VALUE :: ngroups, my_group        ! parai%cp_nogrp, parai%cp_inter_me
       REAL(c_double), INTENT(out) :: rhoe(*)            ! rhoe(nnr1,1)
       REAL(c_double), INTENT(out) :: ekin, rsum_g, rsum_r
       INTEGER(c_int), VALUE :: flags
     END FUNCTION cpb_rhoofr

     ! SUBROUTINE vpsi(c0,c2,f,vpot,psi,nstate,ikind,ispin,redist_c2) vpsi_utils.mod.F90:120
     INTEGER(c_int) FUNCTION cpb_vpsi(plan, c0, c2, ld, nstate, f, vpot, ngroups, my_group, flags) &
          BIND(c, name='cpb_vpsi')
       IMPORT :: c_int, c_long, c_ptr, c_double
       TYPE(c_ptr), VALUE :: plan, c0, c2
       INTEGER(c_long), VALUE :: ld
       INTEGER(c_int), VALUE :: nstate
       REAL(c_double), INTENT(in) :: f(*), vpot(*)
       INTEGER(c_int), VALUE :: ngroups, my_group, flags
     END FUNCTION cpb_vpsi

     ! cntl%tlsd: rhoe(nnr1,2), vpot(nnr1,2), nsup = spin_mod%nsup
     INTEGER(c_int) FUNCTION cpb_rhoofr_lsd(plan, c0, ld_c0, nstate, f, nsup, ngroups, my_group, rhoe, &
          ekin, rsum_g, rsum_r, csums, csumsabs, flags) BIND(c, name='cpb_rhoofr_lsd')
       IMPORT :: c_int, c_long, c_ptr, c_double
       TYPE(c_ptr), VALUE :: plan, c0
       INTEGER(c_long), VALUE :: ld_c0
       INTEGER(c_int), VALUE :: nstate, nsup, ngroups, my_group
       REAL(c_double), INTENT(in) :: f(*)
       REAL(c_double), INTENT(out) :: rhoe(*)
       REAL(c_double), INTENT(out) :: ekin, rsum_g, rsum_r, csums, csumsabs
       INTEGER(c_int), VALUE :: flags
     END FUNCTION cpb_rhoofr_lsd

     INTEGER(c_int) FUNCTION cpb_vpsi_lsd(plan, c0, c2, ld, nstate, f, nsup, vpot, ngroups, my_group, &
          flags) BIND(c, name='cpb_vpsi_lsd')
       IMPORT :: c_int, c_long, c_ptr, c_double
       TYPE(c_ptr), VALUE :: plan, c0, c2
       INTEGER(c_long), VALUE :: ld
       INTEGER(c_int), VALUE :: nstate, nsup, ngroups, my_group, flags
       REAL(c_double), INTENT(in) :: f(*), vpot(*)
     END FUNCTION cpb_vpsi_lsd

     INTEGER(c_int) FUNCTION cpb_c0_invalidate(plan) BIND(c, name='cpb_c0_invalidate')
       IMPORT :: c_int, c_ptr
       TYPE(c_ptr), VALUE :: plan
     END FUNCTION cpb_c0_invalidate

     ! one ikind iteration of rhoofr_c (rhoofr_c_utils.mod.F90:117-178); c0 = C_LOC(c0(1,1,ikind))
     INTEGER(c_int) FUNCTION cpb_rhoofr_kpt(plan, c0, ld, nstate, f, wk, hgkp, hgkm, ngroups, my_group, &
          rhoe, ekin, rsum_g, rsum_r, flags) BIND(c, name='cpb_rhoofr_kpt')
       IMPORT :: c_int, c_long, c_ptr, c_double
       TYPE(c_ptr), VALUE :: plan, c0
       INTEGER(c_long), VALUE :: ld                      ! nkpt%ngwk
       INTEGER(c_int), VALUE :: nstate, ngroups, my_group, flags
       REAL(c_double), INTENT(in) :: f(*), hgkp(*), hgkm(*)   ! crge%f(:,ikk), hgkp(:,ikind), hgkm(:,ikind)
       REAL(c_double), VALUE :: wk                       ! wk(ikk)
       REAL(c_double), INTENT(inout) :: rhoe(*)
       REAL(c_double), INTENT(out) :: ekin, rsum_g, rsum_r
     END FUNCTION cpb_rhoofr_kpt

     ! vpsi with tkpts%tkpnt for k-point ikind (vpsi_utils.mod.F90:562-625)
     INTEGER(c_int) FUNCTION cpb_vpsi_kpt(plan, c0, c2, ld, nstate, f, hgkp, hgkm, vpot, ngroups, &
          my_group, flags) BIND(c, name='cpb_vpsi_kpt')
       IMPORT :: c_int, c_long, c_ptr, c_double
       TYPE(c_ptr), VALUE :: plan, c0, c2
       INTEGER(c_long), VALUE :: ld
       INTEGER(c_int), VALUE :: nstate, ngroups, my_group, flags
       REAL(c_double), INTENT(in) :: f(*), hgkp(*), hgkm(*), vpot(*)
     END FUNCTION cpb_vpsi_kpt

     ! SUBROUTINE tauofr(c0,psi,nstate)  tauofr_utils.mod.F90:42; tau = tauf module array (nnr1,nlsd)
     INTEGER(c_int) FUNCTION cpb_tauofr(plan, c0, ld, nstate, f, nsup, gk, ngroups, my_group, tau, flags) &
          BIND(c, name='cpb_tauofr')
       IMPORT :: c_int, c_long, c_ptr, c_double
       TYPE(c_ptr), VALUE :: plan, c0
       INTEGER(c_long), VALUE :: ld
       INTEGER(c_int), VALUE :: nstate, nsup, ngroups, my_group, flags   ! nsup = -1 without cntl%tlsd
       REAL(c_double), INTENT(in) :: f(*), gk(3,*)
       REAL(c_double), INTENT(out) :: tau(*)
     END FUNCTION cpb_tauofr

     ! SUBROUTINE vtaupsi(c0,c2,f,psi,nstate,ispin)  vtaupsi_utils.mod.F90:38; vtau(nnr1,ispin)
     INTEGER(c_int) FUNCTION cpb_vtaupsi(plan, c0, c2, ld, nstate, f, nsup, gk, vtau, ngroups, my_group, &
          flags) BIND(c, name='cpb_vtaupsi')
       IMPORT :: c_int, c_long, c_ptr, c_double
       TYPE(c_ptr), VALUE :: plan, c0, c2
       INTEGER(c_long), VALUE :: ld
       INTEGER(c_int), VALUE :: nstate, nsup, ngroups, my_group, flags
       REAL(c_double), INTENT(in) :: f(*), gk(3,*), vtau(*)
     END FUNCTION cpb_vtaupsi

     ! vofrhoa :88-102 + vofrhob :155-173 on a plan built from the nhg list; ener(9) = Re/Im eh, ei, ee,
     ! eps and vploc; rhog / vtemp may be C_NULL_PTR; v may be rhoe itself
     INTEGER(c_int) FUNCTION cpb_vofrho_local(plan, rhoe, scg, eivps, eirop, rhog, vtemp, v, ener) &
          BIND(c, name='cpb_vofrho_local')
       IMPORT :: c_int, c_ptr, c_double
       TYPE(c_ptr), VALUE :: plan
       REAL(c_double), INTENT(in) :: rhoe(*), scg(*)
       TYPE(c_ptr), VALUE :: eivps, eirop, rhog, vtemp
       REAL(c_double), INTENT(inout) :: v(*)
       REAL(c_double), INTENT(out) :: ener(9)
     END FUNCTION cpb_vofrho_local

     ! hfx_old(c0,c2,f,psia,nstate,ehfx,vhfx) (hfx_utils.mod.F90:80-965), func1%mhfx = 1, Gamma point, no LSD, no
     ! screening: plan = wavefunction set, plan_dens = pair-density set (nzff / inzf, jhg vectors), scgx(jhg)
     INTEGER(c_int) FUNCTION cpb_hfx(plan, plan_dens, c0, c2, ld, nstate, f, scgx, pfl, ehfx, vhfx, flags) &
          BIND(c, name='cpb_hfx')
       IMPORT :: c_int, c_long, c_ptr, c_double
       TYPE(c_ptr), VALUE :: plan, plan_dens
       TYPE(c_ptr), VALUE :: c0, c2                      ! C_LOC(c0(1,1)), C_LOC(c2(1,1))
       INTEGER(c_long), VALUE :: ld
       INTEGER(c_int), VALUE :: nstate, flags
       REAL(c_double), INTENT(in) :: f(*), scgx(*)
       REAL(c_double), VALUE :: pfl                      ! 0.25 (x func3%phfx for a hybrid)
       REAL(c_double), INTENT(out) :: ehfx, vhfx
     END FUNCTION cpb_hfx

     ! cross-group collectives over NVLink peer memory (device-resident runs, one group per GPU)
     INTEGER(c_int) FUNCTION cpb_peer_create(seg, device, rank, world, bytes, handle_out) &
          BIND(c, name='cpb_peer_create')
       IMPORT :: c_int, c_size_t, c_ptr, c_char
       TYPE(c_ptr), INTENT(out) :: seg
       INTEGER(c_int), VALUE :: device, rank, world      ! rank = parai%cp_inter_me, world = parai%cp_nogrp
       INTEGER(c_size_t), VALUE :: bytes
       CHARACTER(kind=c_char), INTENT(out) :: handle_out(64)
     END FUNCTION cpb_peer_create

     INTEGER(c_int) FUNCTION cpb_peer_connect(seg, all_handles) BIND(c, name='cpb_peer_connect')
       IMPORT :: c_int, c_ptr, c_char
       TYPE(c_ptr), VALUE :: seg
       CHARACTER(kind=c_char), INTENT(in) :: all_handles(64,*)   ! all-gathered over parai%cp_inter_grp
     END FUNCTION cpb_peer_connect

     TYPE(c_ptr) FUNCTION cpb_peer_local_ptr(seg) BIND(c, name='cpb_peer_local_ptr')
       IMPORT :: c_ptr
       TYPE(c_ptr), VALUE :: seg
     END FUNCTION cpb_peer_local_ptr

     INTEGER(c_int) FUNCTION cpb_peer_allreduce_f64(seg, offset, n, stream) BIND(c, name='cpb_peer_allreduce_f64')
       IMPORT :: c_int, c_size_t, c_ptr
       TYPE(c_ptr), VALUE :: seg, stream
       INTEGER(c_size_t), VALUE :: offset, n             ! in doubles, both even
     END FUNCTION cpb_peer_allreduce_f64

     INTEGER(c_int) FUNCTION cpb_peer_bcast_f64(seg, offset, n, src, stream) BIND(c, name='cpb_peer_bcast_f64')
       IMPORT :: c_int, c_size_t, c_ptr
       TYPE(c_ptr), VALUE :: seg, stream
       INTEGER(c_size_t), VALUE :: offset, n
       INTEGER(c_int), VALUE :: src
     END FUNCTION cpb_peer_bcast_f64

     ! cp_grp_redist(C2_vpsi) of a device-resident run (vpsi_utils.mod.F90:708-712): all-gather of the
     ! part_1d state blocks of the (ld, nstate) array that starts `offset` doubles into the segment
     INTEGER(c_int) FUNCTION cpb_peer_redist_c2(seg, offset, ld, nstate, stream) BIND(c, name='cpb_peer_redist_c2')
       IMPORT :: c_int, c_long, c_size_t, c_ptr
       TYPE(c_ptr), VALUE :: seg, stream
       INTEGER(c_size_t), VALUE :: offset
       INTEGER(c_long), VALUE :: ld
       INTEGER(c_int), VALUE :: nstate
     END FUNCTION cpb_peer_redist_c2

     INTEGER(c_int) FUNCTION cpb_peer_allgather_f64(seg, offset, counts, stream) BIND(c, name='cpb_peer_allgather_f64')
       IMPORT :: c_int, c_size_t, c_ptr
       TYPE(c_ptr), VALUE :: seg, stream
       INTEGER(c_size_t), VALUE :: offset
       INTEGER(c_size_t), INTENT(in) :: counts(*)        ! doubles per rank, even
     END FUNCTION cpb_peer_allgather_f64

     ! group-partial ekin / rsum_g / rsum_r summed over the groups (in rank order); synchronises
     INTEGER(c_int) FUNCTION cpb_peer_allreduce_scalars(seg, vals, n, stream) BIND(c, name='cpb_peer_allreduce_scalars')
       IMPORT :: c_int, c_double, c_ptr
       TYPE(c_ptr), VALUE :: seg, stream
       REAL(c_double), INTENT(inout) :: vals(*)
       INTEGER(c_int), VALUE :: n
     END FUNCTION cpb_peer_allreduce_scalars

     INTEGER(c_int) FUNCTION cpb_peer_set_timeout_ms(seg, ms) BIND(c, name='cpb_peer_set_timeout_ms')
       IMPORT :: c_int, c_double, c_ptr
       TYPE(c_ptr), VALUE :: seg
       REAL(c_double), VALUE :: ms
     END FUNCTION cpb_peer_set_timeout_ms

     ! mandatory after the collectives of a step: CPB_ERR_CUDA if a rank missed a barrier
     INTEGER(c_int) FUNCTION cpb_peer_check(seg, stream) BIND(c, name='cpb_peer_check')
       IMPORT :: c_int, c_ptr
       TYPE(c_ptr), VALUE :: seg, stream
     END FUNCTION cpb_peer_check

     INTEGER(c_int) FUNCTION cpb_peer_destroy(seg) BIND(c, name='cpb_peer_destroy')
       IMPORT :: c_int, c_ptr
       TYPE(c_ptr), VALUE :: seg
     END FUNCTION cpb_peer_destroy

     TYPE(c_ptr) FUNCTION cpb_last_error_c() BIND(c, name='cpb_last_error')
       IMPORT :: c_ptr
     END FUNCTION cpb_last_error_c
  END INTERFACE

CONTAINS

  ! the message of the calling thread's last failure as a Fortran string (for stopgm)
  FUNCTION cpb_error_message() RESULT(msg)
    CHARACTER(len=:), ALLOCATABLE :: msg
    TYPE(c_ptr) :: p
    CHARACTER(kind=c_char), POINTER :: s(:)
    INTEGER :: n
    p = cpb_last_error_c()
    msg = ''
    IF (.NOT. c_associated(p)) RETURN
    CALL c_f_pointer(p, s, [1024])
    n = 0
    DO WHILE (n < 1024)
       IF (s(n+1) == c_null_char) EXIT
       n = n + 1
    END DO
    ALLOCATE(CHARACTER(len=n) :: msg)
    msg = TRANSFER(s(1:n), msg)
  END FUNCTION cpb_error_message

END MODULE cpb200_interfaces

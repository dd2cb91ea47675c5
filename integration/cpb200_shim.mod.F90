! cpb200_shim: drop-in guards for CPMD's rhoofr / vpsi that route the supported Gamma-point variants to
! libcpb200.so (include/cpb200.h) and leave every other variant to the original code.
!
! How it is used (integration/rhoofr_vpsi_guard.patch applies exactly this to the reference tree):
!   * add cpb200_interfaces.mod.F90 and this file to src/ (SOURCES list), link with -lcpb200 -lcudart
!   * cpmd.F90, after CALL fft_init:              CALL cpb_shim_init()
!   * rhoofr_utils.mod.F90:122, first statements: CALL cpb_shim_rhoofr(c0,rhoe,psi,nstate,handled); IF (handled) RETURN
!   * vpsi_utils.mod.F90:120,   first statements: CALL cpb_shim_vpsi(c0,c2,f,vpot,psi,nstate,ikind,ispin,redist_c2,handled)
!                                                 IF (handled) RETURN
! The two shim routines take the reference's argument lists unchanged (rhoofr_utils.mod.F90:122-136,
! vpsi_utils.mod.F90:120-135) plus the LOGICAL `handled`.
!
! NOT compile-tested in the authoring image (no Fortran compiler there).  Every module, variable and
! component name below was checked against the reference sources (file:line in the comments); the C
! argument order is checked mechanically against include/cpb200.h by tests/test_abi.py through
! cpb200_interfaces.mod.F90, and tests/test_abi.py::test_fortran_shim_uses_declared_interfaces checks that
! this file only calls entry points that module declares, with the right number of arguments.
MODULE cpb200_shim
  USE, INTRINSIC :: iso_c_binding,     ONLY: c_associated,&
                                             c_double,&
                                             c_int,&
                                             c_loc,&
                                             c_long,&
                                             c_null_ptr,&
                                             c_ptr
  USE cp_cuda_types,                   ONLY: cp_cuda_devices_fft,&   ! cp_cuda_types.mod.F90:26
                                             cp_cuda_env             ! cp_cuda_types.mod.F90:16
  USE cp_grp_utils,                    ONLY: cp_grp_redist           ! cp_grp_utils.mod.F90:23-26
  USE cpb200_interfaces,               ONLY: CPB_C0_KEEP,&
                                             CPB_C0_REUSE,&
                                             CPB_VPSI_TKSHAM,&
                                             cpb_error_message,&
                                             cpb_length_supported,&
                                             cpb_plan_create,&
                                             cpb_plan_destroy,&
                                             cpb_rhoofr,&
                                             cpb_rhoofr_lsd,&
                                             cpb_vpsi,&
                                             cpb_vpsi_lsd
  USE cppt,                            ONLY: hg,&                    ! cppt.mod.F90:45
                                             inyh                    ! cppt.mod.F90:26
  USE dg,                              ONLY: tdgcomm                 ! dg.mod.F90:26
  USE elct,                            ONLY: crge                    ! rhoofr_utils.mod.F90:38
  USE ener,                            ONLY: chrg,&                  ! ener.mod.F90:117-120
                                             ener_com                ! ener.mod.F90:50
  USE error_handling,                  ONLY: stopgm
  USE kinds,                           ONLY: real_8
  USE kpts,                            ONLY: tkpts                   ! vpsi_utils.mod.F90:54
  USE mp_interface,                    ONLY: mp_sum                  ! mp_interface.mod.F90:170
  USE parac,                           ONLY: parai,&                 ! parac.mod.F90:42-58
                                             paral
  USE prcp,                            ONLY: prcp_com                ! prcp.mod.F90:35
  USE pslo,                            ONLY: pslo_com                ! pslo.mod.F90:19
  USE rswfmod,                         ONLY: rsactive                ! rswf.mod.F90
  USE spin,                            ONLY: clsd,&
                                             lspin2,&                ! spin.mod.F90:55
                                             spin_mod                ! spin.mod.F90:34
  USE system,                          ONLY: cntl,&
                                             fpar,&
                                             group,&
                                             locpot2,&               ! system.mod.F90:830
                                             ncpw,&
                                             parm,&                  ! system.mod.F90:143-145
                                             spar
  USE td_input,                        ONLY: td_prop                 ! td_input.mod.F90:73

  IMPLICIT NONE

  PRIVATE

  PUBLIC :: cpb_shim_init
  PUBLIC :: cpb_shim_finalize
  PUBLIC :: cpb_shim_rhoofr
  PUBLIC :: cpb_shim_vpsi

  TYPE(c_ptr), SAVE                          :: cpb_plan = c_null_ptr

CONTAINS

  ! ==================================================================
  SUBROUTINE cpb_shim_init()
    ! ==--------------------------------------------------------------==
    ! == Creates the plan from the module globals fft_init has set    ==
    ! == (SURVEY.md 8b).  Eligible runs: GPU FFT requested, one MPI   ==
    ! == rank per state group (CP_GROUPS = ranks = GPUs, so that      ==
    ! == fpar%kr1 == fpar%kr1s and the plane-wave distribution of     ==
    ! == loadpa collapses), all three mesh lengths instantiated.      ==
    ! ==--------------------------------------------------------------==
    CHARACTER(*), PARAMETER                  :: procedureN = 'cpb_shim_init'

    INTEGER(c_int)                           :: device_idx, ierr, kr(3), &
                                                nr(3)

    IF (c_associated(cpb_plan)) RETURN
    IF (.NOT.cp_cuda_env%use_fft) RETURN
    IF (parai%nproc.NE.1) RETURN            ! ranks inside one state group
    IF (group%nogrp.GT.1) RETURN
    nr = [spar%nr1s, spar%nr2s, spar%nr3s]
    kr = [fpar%kr1, fpar%kr2s, fpar%kr3s]
    IF (cpb_length_supported(nr(1)).NE.1 .OR. cpb_length_supported(nr(2)).NE.1 .OR. &
         cpb_length_supported(nr(3)).NE.1) RETURN
    device_idx = 0
    IF (ALLOCATED(cp_cuda_devices_fft%ids)) device_idx = cp_cuda_devices_fft%ids(1)
    ! 0: batch size chosen by the library from the mesh
    ierr = cpb_plan_create(cpb_plan, nr, kr, ncpw%ngw, inyh, hg, parm%tpiba2, parm%omega, &
         device_idx, 0_c_int)
    IF (ierr.NE.0) CALL stopgm(procedureN, cpb_error_message(), __LINE__, __FILE__)
    ! ==--------------------------------------------------------------==
    RETURN
  END SUBROUTINE cpb_shim_init
  ! ==================================================================
  SUBROUTINE cpb_shim_finalize()
    INTEGER(c_int)                           :: ierr

    IF (c_associated(cpb_plan)) ierr = cpb_plan_destroy(cpb_plan)
    cpb_plan = c_null_ptr
    RETURN
  END SUBROUTINE cpb_shim_finalize
  ! ==================================================================
  SUBROUTINE cpb_shim_rhoofr(c0,rhoe,psi,nstate,handled)
    ! ==--------------------------------------------------------------==
    ! == Same arguments as rhoofr (rhoofr_utils.mod.F90:122-136).     ==
    ! == handled = .TRUE.: rhoe, ener_com%ekin, chrg%csumg/csumr (and ==
    ! == csums/csumsabs with LSD) are set like the original would;    ==
    ! == handled = .FALSE.: nothing was touched, run the original.    ==
    ! ==--------------------------------------------------------------==
    COMPLEX(real_8), TARGET                  :: c0(:,:)
    REAL(real_8), TARGET __CONTIGUOUS        :: rhoe(:,:)
    COMPLEX(real_8), TARGET __CONTIGUOUS     :: psi(:)
    INTEGER                                  :: nstate
    LOGICAL                                  :: handled

    CHARACTER(*), PARAMETER                  :: procedureN = 'cpb_shim_rhoofr'
    REAL(real_8), PARAMETER                  :: delta = 1.e-6_real_8    ! rhoofr_utils.mod.F90:140

    INTEGER                                  :: i
    INTEGER(c_int)                           :: ierr
    REAL(real_8)                             :: buf(2), rsum1, rsum1abs
    REAL(c_double)                           :: csums, csumsabs, ekin_blk, &
                                                rsum_blk, rsumr_blk

    handled = .FALSE.
    IF (.NOT.c_associated(cpb_plan)) RETURN
    ! variants the library does not implement take the original path
    ! (rhoofr_utils.mod.F90:187 tdg, :350 rsactive, :386 LSE, :498 Vanderbilt, :637 tau is handled below)
    IF (tkpts%tkpnt .OR. lspin2%tlse .OR. tdgcomm%tdg .OR. rsactive .OR. pslo_com%tivan .OR. cntl%cdft) RETURN
    IF (parai%nproc.NE.1 .OR. group%nogrp.GT.1) RETURN
    IF (cntl%ttau) RETURN                   ! tauofr follows in the original (:637): keep it there
    IF (.NOT.IS_CONTIGUOUS(c0)) RETURN

    IF (cntl%tlsd) THEN
       ! rhoe(:,1) = alpha+beta, rhoe(:,2) = beta (rhoofr_utils.mod.F90:543-559)
       ierr = cpb_rhoofr_lsd(cpb_plan, c_loc(c0(1,1)), INT(SIZE(c0,1),c_long), INT(nstate,c_int), crge%f(:,1), &
            INT(spin_mod%nsup,c_int), INT(parai%cp_nogrp,c_int), INT(parai%cp_inter_me,c_int), rhoe, &
            ekin_blk, rsum_blk, rsumr_blk, csums, csumsabs, CPB_C0_KEEP)
    ELSE
       ierr = cpb_rhoofr(cpb_plan, c_loc(c0(1,1)), INT(SIZE(c0,1),c_long), INT(nstate,c_int), crge%f(:,1), &
            INT(parai%cp_nogrp,c_int), INT(parai%cp_inter_me,c_int), rhoe, ekin_blk, rsum_blk, rsumr_blk, &
            CPB_C0_KEEP)
    ENDIF
    IF (ierr.NE.0) CALL stopgm(procedureN, cpb_error_message(), __LINE__, __FILE__)

    ! The library returns the sums of the calling group's block of states; the original computes ekin and
    ! rsum over ALL states on every group (kin_energy, rhoofr_utils.mod.F90:178) and sums rhoe over the
    ! groups (:457-461)
    IF (parai%cp_nogrp.GT.1) THEN
       IF (cntl%tlsd) THEN
          ! the library returned the group's partial alpha and beta densities: sum them over the groups,
          ! then form alpha+beta / beta and the spin sums like rhoofr_utils.mod.F90:543-559
          CALL cp_grp_redist(rhoe, fpar%nnr1, 2)
          rsum1 = 0._real_8
          rsum1abs = 0._real_8
          DO i = 1, fpar%nnr1
             rsum1 = rsum1 + (rhoe(i,1) - rhoe(i,2))
             rsum1abs = rsum1abs + ABS(rhoe(i,1) - rhoe(i,2))
             rhoe(i,1) = rhoe(i,1) + rhoe(i,2)
          ENDDO
          csums = rsum1*parm%omega/REAL(spar%nr1s*spar%nr2s*spar%nr3s,kind=real_8)
          csumsabs = rsum1abs*parm%omega/REAL(spar%nr1s*spar%nr2s*spar%nr3s,kind=real_8)
       ELSE
          CALL cp_grp_redist(rhoe, fpar%nnr1, clsd%nlsd)
       ENDIF
       buf(1) = ekin_blk
       buf(2) = rsum_blk
       CALL mp_sum(buf, 2, parai%cp_inter_grp)
       ekin_blk = buf(1)
       rsum_blk = buf(2)
    ENDIF
    ener_com%ekin = ekin_blk                ! kin_energy_utils.mod.F90:110
    ! charge check of rhoofr_utils.mod.F90:603-635 (allgrp has one member here)
    rsum1 = 0._real_8
    DO i = 1, fpar%nnr1
       rsum1 = rsum1 + rhoe(i,1)
    ENDDO
    chrg%csumg = rsum_blk
    chrg%csumr = rsum1*parm%omega/REAL(spar%nr1s*spar%nr2s*spar%nr3s,kind=real_8)
    IF (cntl%tlsd) THEN
       chrg%csums = csums
       chrg%csumsabs = csumsabs
    ENDIF
    IF (paral%parent .AND. ABS(chrg%csumr-chrg%csumg).GT.delta) THEN
       IF (paral%io_parent) WRITE(6,'(A,T46,F20.12)') ' IN FOURIER SPACE:', chrg%csumg
       IF (paral%io_parent) WRITE(6,'(A,T46,F20.12)') ' IN REAL SPACE:', chrg%csumr
       CALL stopgm(procedureN, 'TOTAL DENSITY SUMS ARE NOT EQUAL', __LINE__, __FILE__)
    ENDIF
    handled = .TRUE.
    ! ==--------------------------------------------------------------==
    RETURN
  END SUBROUTINE cpb_shim_rhoofr
  ! ==================================================================
  SUBROUTINE cpb_shim_vpsi(c0,c2,f,vpot,psi,nstate,ikind,ispin,redist_c2,handled)
    ! ==--------------------------------------------------------------==
    ! == Same arguments as vpsi (vpsi_utils.mod.F90:120-135).         ==
    ! == handled = .TRUE.: c2 += -f/2 (1/2 G^2 c0 + FFT[V psi]) for   ==
    ! == the calling group's block of states (add_wfn, :717), summed  ==
    ! == over the groups when redist_c2 (:708-712).                   ==
    ! ==--------------------------------------------------------------==
    COMPLEX(real_8), TARGET                  :: c0(:,:), c2(:,:)
    REAL(real_8)                             :: f(:)
    REAL(real_8), TARGET __CONTIGUOUS        :: vpot(:,:)
    COMPLEX(real_8), TARGET __CONTIGUOUS     :: psi(:)
    INTEGER                                  :: nstate, ikind, ispin
    LOGICAL                                  :: redist_c2, handled

    CHARACTER(*), PARAMETER                  :: procedureN = 'cpb_shim_vpsi'

    INTEGER(c_int)                           :: flags, ierr

    handled = .FALSE.
    IF (.NOT.c_associated(cpb_plan)) RETURN
    IF (ikind.NE.1) RETURN
    ! vpsi_utils.mod.F90:176 tdg, :238 k-points, :239 LSE, :245 cdft, :487-561 external / local potentials,
    ! :641-647 akin
    IF (tkpts%tkpnt .OR. lspin2%tlse .OR. tdgcomm%tdg .OR. cntl%cdft .OR. cntl%ttau .OR. rsactive) RETURN
    IF (td_prop%td_extpot .OR. locpot2%tlpot .OR. prcp_com%akin.GT.1.e-10_real_8) RETURN
    IF (parai%nproc.NE.1 .OR. group%nogrp.GT.1) RETURN
    IF (SIZE(c0,1).NE.SIZE(c2,1)) RETURN
    IF (.NOT.(IS_CONTIGUOUS(c0) .AND. IS_CONTIGUOUS(c2))) RETURN
    IF (cntl%tlsd .AND. ispin.NE.2) RETURN  ! LSD diagonalisation schemes that pass one spin channel
    IF (.NOT.cntl%tlsd .AND. ispin.NE.1) RETURN

    flags = CPB_C0_REUSE                    ! the block rhoofr uploaded in this step, if it is still valid
    IF (cntl%tksham) flags = IOR(flags, CPB_VPSI_TKSHAM)   ! vpsi_utils.mod.F90:628-633
    IF (cntl%tlsd) THEN
       ierr = cpb_vpsi_lsd(cpb_plan, c_loc(c0(1,1)), c_loc(c2(1,1)), INT(SIZE(c0,1),c_long), INT(nstate,c_int), f, &
            INT(spin_mod%nsup,c_int), vpot, INT(parai%cp_nogrp,c_int), INT(parai%cp_inter_me,c_int), flags)
    ELSE
       ierr = cpb_vpsi(cpb_plan, c_loc(c0(1,1)), c_loc(c2(1,1)), INT(SIZE(c0,1),c_long), INT(nstate,c_int), f, &
            vpot, INT(parai%cp_nogrp,c_int), INT(parai%cp_inter_me,c_int), flags)
    ENDIF
    IF (ierr.NE.0) CALL stopgm(procedureN, cpb_error_message(), __LINE__, __FILE__)
    ! The original sums the temporary C2_vpsi over the groups before add_wfn (vpsi_utils.mod.F90:708-717); summing c2
    ! itself is the same thing when c2 is zero outside the group's block on entry, which is how the callers
    ! that pass redist_c2=.TRUE. use it (forces_driver.mod.F90:175,224,283: zeroing(c2), vpsi, cp_grp_redist)
    IF (redist_c2 .AND. parai%cp_nogrp.GT.1) CALL cp_grp_redist(c2, SIZE(c2,1), nstate)
    handled = .TRUE.
    ! ==--------------------------------------------------------------==
    RETURN
  END SUBROUTINE cpb_shim_vpsi
  ! ==================================================================

END MODULE cpb200_shim

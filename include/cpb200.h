/* cpb200 — B200-native (sm_100a) drop-in for CPMD's Gamma-point `vpsi` + `rhoofr` hot path.
 *
 * C ABI: plain pointers and sizes, no C++/torch types.  All entry points return 0 on success or
 * a negative cpb_status; cpb_last_error() gives the message of the calling thread's last failure
 * (the Fortran shim turns a non-zero return into CALL stopgm(...), the reference's only error
 * convention, error_handling.mod.F90:11-53).
 *
 * What each entry point replaces in the reference (paths relative to /root/reference/src):
 *
 *   cpb_plan_create   the module-global FFT/G-vector state built by fft_init/fftprp_default_init
 *                     (fftprp_utils.mod.F90:66-93,139-285: mg ray table, kr3min/kr3max, nzhs,
 *                     indzs, msp) and setfftn(0) (fftnew_utils.mod.F90:53-175), gathered once into
 *                     a plan.  Inputs are exactly the globals the Fortran shim can read:
 *                     spar%nr1s.., fpar%kr1.., ncpw%ngw, inyh(3,ngw), hg(ngw) (cppt.mod.F90:22-45),
 *                     parm%tpiba2, parm%omega (system.mod.F90:138-151).
 *   cpb_rhoofr[_dev]  SUBROUTINE rhoofr(c0,rhoe,psi,nstate)        rhoofr_utils.mod.F90:122-644
 *                     incl. kin_energy (kin_energy_utils.mod.F90:62-110) and the charge sums
 *                     (rhoofr_utils.mod.F90:607-619).  `psi` (scratch) has no counterpart: the
 *                     library owns its work space.
 *   cpb_vpsi[_dev]    SUBROUTINE vpsi(c0,c2,f,vpot,psi,nstate,ikind,ispin,redist_c2)
 *                                                                  vpsi_utils.mod.F90:120-732
 *   cpb_part_1d_*     part_1d_nbr_el_in_blk / part_1d_get_el_in_blk   part_1d.mod.F90:22-57
 *
 * State groups (CP_GROUPS, set_cp_grp_utils.mod.F90:77-83): every call takes (ngroups, my_group)
 * and processes only the group's contiguous block of states, pairing states inside the block
 * exactly like the reference loops (vpsi_utils.mod.F90:376-383, rhoofr_utils.mod.F90:306-310).
 * The cross-group reductions stay with the caller, where the reference has them:
 * cp_grp_redist(rhoe) (rhoofr_utils.mod.F90:457-461) and, if wanted, cp_grp_redist(c2)
 * (vpsi_utils.mod.F90:708-712 / forces_driver.mod.F90:283).  Deviation, documented in
 * ekin / rsum_g are returned for the group's block only (the reference computes
 * them redundantly over all states on every group) — sum them over groups together with rhoe.
 *
 * Layout conventions are the reference's: c0/c2 are column-major (ld, nstate) COMPLEX*16, one
 * column per state, ngw <= ld; rhoe/vpot are REAL*8 (kr1, kr2s, kr3s) x-fastest with the
 * odd-padded leading dimensions of leadim (loadpa_utils.mod.F90:509-525); inyh is INTEGER*4
 * (3,ngw), 1-based.  Implemented variants: Gamma point RKS and LSD (cpb_rhoofr*, cpb_vpsi*), one
 * k-point at a time without LSD (cpb_*_kpt_dev), meta-GGA tau (cpb_tauofr_dev, cpb_vtaupsi_dev), the
 * local part of vofrho between the two routines (cpb_vofrho_local*); no LSE, no double grid, akin = 0
 * — the shim must route everything else to the original routine (list in INTEGRATION.md).
 */
#ifndef CPB200_H
#define CPB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct cpb_plan cpb_plan;

typedef enum {
  CPB_OK = 0,
  CPB_ERR_INVALID = -1,     /* bad argument / unsupported mesh length / inconsistent G list */
  CPB_ERR_CUDA = -2,        /* CUDA runtime failure */
  CPB_ERR_NOMEM = -3,
  CPB_ERR_UNSUPPORTED = -4, /* requested variant is outside the implemented path */
  CPB_ERR_CHARGE = -5       /* |rsum_r - rsum_g| > 1e-6 (rhoofr_utils.mod.F90:625-635), opt-in */
} cpb_status;

/* flags for cpb_vpsi / cpb_rhoofr */
#define CPB_VPSI_OVERWRITE 1u /* c2 = result instead of the reference's c2 += result */
#define CPB_VPSI_TKSHAM 2u    /* cntl%tksham: f==0 -> fi=0.5 instead of 1 (vpsi_utils:628-633) */
#define CPB_RHO_CHECK_CHARGE 1u /* return CPB_ERR_CHARGE like the reference's stopgm */
#define CPB_RHO_ACCUMULATE 2u   /* cpb_rhoofr_kpt*: add to rhoe instead of zeroing it first (k-points after the first) */
/* host-pointer entry points only: device-side cache of the group's c0 block */
#define CPB_C0_KEEP 0x10u  /* after this call the uploaded block stays valid on the device */
#define CPB_C0_REUSE 0x20u /* skip the upload if (c0 pointer, ld, nstate, group) match the kept block */
/* all entry points: device-side counterpart of CPMD's REAL SPACE WFN KEEP (rsactive: rhoofr stores
 * the real-space wavefunctions in rswf, rhoofr_utils.mod.F90:350-363, so that later FFTs of the same
 * orbitals can be skipped).  cpb_rhoofr* with CPB_PSI_KEEP leaves the y-pass output of every pair of
 * the block (16*n1*n2*zband bytes per pair, half the size of psi(r)) in HBM; the next cpb_vpsi* call
 * with CPB_PSI_REUSE and the same (c0 pointer, ld, nstate, group, nsup) starts from it - no gather, no
 * x and y inverse passes - and consumes it.  The caller vouches that c0 did not change in between
 * (MD step: forces_driver.mod.F90:165 -> :224).  Silently ignored when the device has no room. */
#define CPB_PSI_KEEP 0x40u
#define CPB_PSI_REUSE 0x80u
/* device-pointer entry points cpb_rhoofr[_lsd]_dev / cpb_vpsi[_lsd]_dev: only ENQUEUE the work on `stream` and
 * return (like every CUDA library call; later work on the stream sees the results).  cpb_vpsi*_dev has nothing
 * to hand back.  cpb_rhoofr*_dev leaves ekin / rsum_g / rsum_r (/ csums / csumsabs) untouched: fetch them with
 * cpb_rhoofr_finish, which waits for the partial sums of that call only - work enqueued behind it (the group
 * exchanges, the vpsi of the same step) keeps running.  One pending rhoofr per plan.  `f` is copied.  Ignored
 * (the call synchronises as usual) while per-kernel profiling is on. */
#define CPB_ASYNC 0x100u

typedef struct {
  int nr[3];          /* mesh */
  int kr[3];          /* padded leading dimensions */
  int ngw;
  int geq0;           /* first plane wave is G=0 (derived from inyh) */
  int nrays;          /* rays transformed by the x pass (== msrays for a convex cutoff) */
  int zband;          /* kr3max-kr3min+1 */
  int xband;          /* x extent holding coefficients */
  int max_batch;      /* packed pairs per kernel batch */
  int device;
  int radix[3][2];    /* two-pass factorisation of each axis */
  size_t workspace_bytes;
  int band_pruned[3]; /* 1: the axis runs the band-pruned kernel variants (coefficients in the middle half) */
  int chunk_xtiles;   /* x tiles (of 8 columns) per y/z chunk (default: all of them, one y and one z launch per batch) */
  int streams;        /* work spaces / streams the batches of a call alternate between */
  int z_warp_kernels; /* 1: the z passes of rhoofr / vpsi run the warp-autonomous kernels (kernels_zw.h) */
  int z_warp_radix[2];/* their factorisation of n3 (band-side radix, real-space-side radix), 0 if unused */
  int x_warp_kernels; /* bit 0: the inverse x pass runs the warp-autonomous kernel (kernels_xw.h), bit 1: the forward one */
  int x_warp_radix;   /* their band-side radix of n1 (the other one is 8), 0 if unused */
} cpb_plan_info;

const char* cpb_last_error(void);
const char* cpb_version(void);

/* 1 if mesh length n has a kernel instantiation in this build, else 0 */
int cpb_length_supported(int n);

/* nr, kr: 3 ints each.  inyh: (3,ngw) column-major INTEGER*4, 1-based (cppt inyh).  hg: |G|^2 in
 * units of tpiba2.  device: CUDA ordinal.  max_batch_pairs: pairs per batch (<=0: sized
 * from the mesh - about 3 GB of work space, a multiple of 8 in [8, 64]; 32 for the 192^3 north-star case). */
int cpb_plan_create(cpb_plan** plan, const int* nr, const int* kr, int ngw, const int32_t* inyh,
                    const double* hg, double tpiba2, double omega, int device,
                    int max_batch_pairs);
int cpb_plan_destroy(cpb_plan* plan);
int cpb_plan_get_info(const cpb_plan* plan, cpb_plan_info* info);

/* Reference-compatible index maps derived by the plan (for cross-checking against the host
 * program's own nzhs/indzs, fftprp_utils.mod.F90:269-285).  Both arrays have ngw entries, 1-based
 * into (kr1s, nrays) ray storage with the reference's ray numbering. */
int cpb_plan_get_maps(const cpb_plan* plan, int32_t* nzhs, int32_t* indzs);

/* part_1d.mod.F90:22-57 (proc is 0-based, elements 1-based, as in the reference) */
int cpb_part_1d_nbr_el_in_blk(int n_elem, int proc, int nproc);
int cpb_part_1d_get_el_in_blk(int i_elem, int n_elem, int proc, int nproc);

/* ---- host-pointer entry points: what the Fortran shim binds --------------------------------
 * Arrays live in host memory; the library stages them (pinned bounce buffers, copies overlapped
 * with the kernels of the previous batch) and copies the results back before returning. */
int cpb_rhoofr(cpb_plan* plan, const void* c0, long ld_c0, int nstate, const double* f,
               int ngroups, int my_group, double* rhoe, double* ekin, double* rsum_g,
               double* rsum_r, unsigned flags);

int cpb_vpsi(cpb_plan* plan, const void* c0, void* c2, long ld, int nstate, const double* f,
             const double* vpot, int ngroups, int my_group, unsigned flags);

/* ---- LSD (cntl%tlsd) variants --------------------------------------------------------------
 * States 1..nsup are alpha, nsup+1..nstate beta (spin_mod%nsup).  rhoe and vpot have two columns,
 * (nnr1, 2) column-major like the reference's rhoe(nnr1,nlsd) / vpot(nnr1,ispin):
 *   cpb_rhoofr_lsd  rhoofr with cntl%tlsd (rhoofr_utils.mod.F90:375-385, 543-559).  ngroups == 1:
 *                   on return column 1 = alpha+beta density, column 2 = beta density, csums =
 *                   integral of alpha-beta, csumsabs = integral of |alpha-beta| (chrg%csums,
 *                   chrg%csumsabs).  ngroups > 1: the columns hold the group's partial alpha and
 *                   beta densities - the reference applies cp_grp_redist(rhoe,nnr1,nlsd) (:457-461)
 *                   before that step - rsum_r / csums are the group's partial sums and csumsabs is
 *                   not meaningful; finish with the original lines :543-559 or cpb_lsd_finish_dev.
 *   cpb_vpsi_lsd    vpsi with cntl%tlsd and ispin = 2 (vpsi_utils.mod.F90:450-482): column 1 of vpot
 *                   acts on the alpha states, column 2 on the beta states.
 * Every launch works on one spin channel; the one pair that straddles the spin boundary is
 * transformed as two single states (the same linear map as the reference's mixed pair). */
int cpb_rhoofr_lsd(cpb_plan* plan, const void* c0, long ld_c0, int nstate, const double* f, int nsup,
                   int ngroups, int my_group, double* rhoe, double* ekin, double* rsum_g,
                   double* rsum_r, double* csums, double* csumsabs, unsigned flags);
int cpb_vpsi_lsd(cpb_plan* plan, const void* c0, void* c2, long ld, int nstate, const double* f,
                 int nsup, const double* vpot, int ngroups, int my_group, unsigned flags);
int cpb_rhoofr_lsd_dev(cpb_plan* plan, const void* c0_dev, long ld_c0, int nstate, const double* f,
                       int nsup, int ngroups, int my_group, double* rhoe_dev, double* ekin,
                       double* rsum_g, double* rsum_r, double* csums, double* csumsabs,
                       unsigned flags, void* stream);
int cpb_vpsi_lsd_dev(cpb_plan* plan, const void* c0_dev, void* c2_dev, long ld, int nstate,
                     const double* f, int nsup, const double* vpot_dev, int ngroups, int my_group,
                     unsigned flags, void* stream);
/* rhoofr_utils.mod.F90:543-559 on the group-summed channel densities (device array (nnr1,2)):
 * column 1 += column 2, and the three integrals. */
int cpb_lsd_finish_dev(cpb_plan* plan, double* rhoe_dev, double* rsum_r, double* csums,
                       double* csumsabs, void* stream);

/* Optional: keep the group's block of c0 on the device between rhoofr and the following vpsi of
 * the same MD step (the reference's cp_cuwfn cache, vpsi_utils.mod.F90:268-273, which trusts a
 * checksum of the first state; here the caller states it explicitly).  After cpb_c0_upload, or a
 * host call made with CPB_C0_KEEP, host entry points called with CPB_C0_REUSE and the same c0
 * pointer, ld, nstate and group reuse the device copy; cpb_c0_invalidate drops it. */
int cpb_c0_upload(cpb_plan* plan, const void* c0, long ld_c0, int nstate, int ngroups, int my_group);
int cpb_c0_invalidate(cpb_plan* plan);

/* ---- device-pointer entry points (benchmark / Python harness, multi-GPU driver) ------------
 * c0/c2/rhoe/vpot are device pointers on the plan's device; f is a host array; stream is a
 * cudaStream_t (NULL = default stream).  The call returns after the stream work completed
 * (scalars are written to host memory). */
int cpb_rhoofr_dev(cpb_plan* plan, const void* c0_dev, long ld_c0, int nstate, const double* f,
                   int ngroups, int my_group, double* rhoe_dev, double* ekin, double* rsum_g,
                   double* rsum_r, unsigned flags, void* stream);

int cpb_vpsi_dev(cpb_plan* plan, const void* c0_dev, void* c2_dev, long ld, int nstate,
                 const double* f, const double* vpot_dev, int ngroups, int my_group,
                 unsigned flags, void* stream);

/* One-shot ordering hint for the next cpb_vpsi*_dev call: `event` (a cudaEvent_t the caller recorded on
 * the stream that PRODUCES vpot, e.g. the stream of cpb_peer_bcast_f64 or of the caller's vofrho) is
 * waited for only before the first kernel that reads vpot - the z pass of the first batch - so the
 * gather and the x / y passes of that batch overlap the producer instead of queueing behind it.  Without
 * it the caller orders the whole call after the producer as usual.  NULL clears a pending hint. */
int cpb_plan_set_vpot_event(cpb_plan* plan, void* event);

/* Second half of a cpb_rhoofr_dev / cpb_rhoofr_lsd_dev call made with CPB_ASYNC: waits until that call's
 * partial sums have reached the host and returns what the synchronous call returns (csums / csumsabs: LSD
 * only, may be NULL).  CPB_RHO_CHECK_CHARGE of the call is honoured here. */
int cpb_rhoofr_finish(cpb_plan* plan, double* ekin, double* rsum_g, double* rsum_r, double* csums, double* csumsabs);
/* 1 while a CPB_ASYNC rhoofr waits for its cpb_rhoofr_finish (0 if the call synchronised after all: profiling on) */
int cpb_rhoofr_pending(cpb_plan* plan);

/* ---- dense transforms on the density cutoff and the local part of vofrho -------------------
 * (the step between rhoofr and vpsi.)  These run on a plan created by cpb_plan_create
 * from the nhg vectors of the DENSITY cutoff: ngw := ncpw%nhg, inyh(3,nhg), hg(nhg) - the same
 * cppt arrays, whose first ngw entries are the wavefunction sphere - so that the plan's maps are
 * the reference's nzh / indz (fftprp_utils.mod.F90:269-285) instead of nzhs / indzs.  Real-space
 * arrays are REAL*8 (nnr1, nfields), G-space arrays COMPLEX*16 (ld, nfields), column-major.
 *
 *   cpb_dense_fwfft_dev   v = CMPLX(f, 0); CALL fwfftn(v,.FALSE.); g(ig) = v(nzh(ig))
 *                         (vofrhoa_utils.mod.F90:88-95 + ppener_utils.mod.F90:91 / zgthr in
 *                         vofrhob_utils.mod.F90:240,245; transform: fftmain_utils.mod.F90:137-153 with
 *                         phasen, fftutil_utils.mod.F90:479-503).  nfields = 2 packs two real fields
 *                         into one complex transform and separates them at +-G.
 *   cpb_dense_invfft_dev  v(nzh) = g, v(indz) = CONJG(g); CALL invfftn(v,.FALSE.); f = REAL(v)
 *                         (vofrhob_utils.mod.F90:155-173, fftmain_utils.mod.F90:105-120).  nfields = 2:
 *                         v = g1 + i g2 at +G, conj(g1) + i conj(g2) at -G; f(:,1) = REAL(v),
 *                         f(:,2) = AIMAG(v).  CPB_DENSE_ACCUMULATE: f += instead of f =.
 *   cpb_vofrho_local[_dev]  the G-space electrostatics between the two: rhog = FFT(rhoe(:,1));
 *                         ppener (ppener_utils.mod.F90:23-108): vtemp = scg*(rhog+eirop)+eivps and
 *                         the sums eh, ei, ee, eps; v = REAL(FFT^-1(vtemp)).  Inputs scg (cppt scg),
 *                         eivps / eirop (eicalc) have nhg entries.  ener[9] (host) = Re eh, Im eh,
 *                         Re ei, Im ei, Re ee, Im ee, Re eps, Im eps, vploc - the caller forms
 *                         ehep = Re(eh)*omega, epseu = 2*Re(eps)*omega ... (vofrhoa_utils.mod.F90:104-127).
 *                         rhog / vtemp may be NULL (not returned).  v may alias rhoe, like the
 *                         reference's rhoe (in: density, out: potential).  Exchange-correlation
 *                         (xcener, gcener) is NOT included: the caller adds it to v.
 */
#define CPB_DENSE_ACCUMULATE 1u
int cpb_dense_fwfft_dev(cpb_plan* plan, const double* f_dev, int nfields, void* g_dev, long ld, void* stream);
int cpb_dense_invfft_dev(cpb_plan* plan, const void* g_dev, long ld, int nfields, double* f_dev,
                         unsigned flags, void* stream);
int cpb_vofrho_local_dev(cpb_plan* plan, const double* rhoe_dev, const double* scg_dev,
                         const void* eivps_dev, const void* eirop_dev, void* rhog_dev, void* vtemp_dev,
                         double* v_dev, double* ener, void* stream);
int cpb_vofrho_local(cpb_plan* plan, const double* rhoe, const double* scg, const void* eivps,
                     const void* eirop, void* rhog, void* vtemp, double* v, double* ener);

/* ---- k-points (tkpts%tkpnt) ------------------------------------------------------------------
 * Complex states, one per transform (njump = 1, vpsi_utils.mod.F90:237-238).  One k-point per call,
 * the caller keeps the reference's loops over k-point blocks (rhoofr_c_utils.mod.F90:112-116; vpsi's
 * ikind argument).  c0/c2 are (ld >= nkpt%ngwk = 2*ngw, nstate): components 1..ngw belong to +G,
 * ngw+1..2*ngw to -G (set_psi_1_state_g_kpts, state_utils.mod.F90:192-224).  hgkp/hgkm: |k+G|^2,
 * |k-G|^2 of this k-point, ngw doubles each (kpts hgkp(:,ikind), hgkm(:,ikind)).
 *   cpb_rhoofr_kpt_dev  one ikind iteration of rhoofr_c (rhoofr_c_utils.mod.F90:117-178): rhoe +=
 *                       wk f_i / omega |psi_i(r)|^2 over the group's block; f = crge%f(:,ikk), wk =
 *                       wk(ikk).  Without CPB_RHO_ACCUMULATE rhoe is zeroed first (:107, first k-point).
 *                       ekin / rsum_g: this k-point's (and block's) contribution to ener_com%ekin
 *                       (:138,182) and chrg%csumg (:119); rsum_r: integral of rhoe as it stands.
 *   cpb_vpsi_kpt_dev    vpsi with tkpts%tkpnt for k-point ikind (vpsi_utils.mod.F90:432,487-493,
 *                       562-564,614-625): fi = f (2 if zero), c2(ig) += -fi (tpiba2/2 hgkp c0(ig) +
 *                       FFT[V psi](+G)), c2(ig+ngw) += -fi (tpiba2/2 hgkm c0(ig+ngw) + FFT[V psi](-G)),
 *                       the -G slot of G=0 is left alone (zeroed with CPB_VPSI_OVERWRITE).
 * Not covered (the shim takes the original path): k-points with cntl%tlsd (the reference applies the
 * Gamma-only mixed-pair formula to state nsup, vpsi_utils.mod.F90:451-469), tgaugep/tgaugef, tkblock
 * swapping (the caller's job), Vanderbilt, symrho. */
int cpb_rhoofr_kpt_dev(cpb_plan* plan, const void* c0_dev, long ld, int nstate, const double* f, double wk,
                       const double* hgkp_dev, const double* hgkm_dev, int ngroups, int my_group,
                       double* rhoe_dev, double* ekin, double* rsum_g, double* rsum_r, unsigned flags,
                       void* stream);
int cpb_vpsi_kpt_dev(cpb_plan* plan, const void* c0_dev, void* c2_dev, long ld, int nstate, const double* f,
                     const double* hgkp_dev, const double* hgkm_dev, const double* vpot_dev, int ngroups,
                     int my_group, unsigned flags, void* stream);
/* host-pointer forms (what the Fortran shim binds): all arrays in host memory, staged per call */
int cpb_rhoofr_kpt(cpb_plan* plan, const void* c0, long ld, int nstate, const double* f, double wk,
                   const double* hgkp, const double* hgkm, int ngroups, int my_group, double* rhoe,
                   double* ekin, double* rsum_g, double* rsum_r, unsigned flags);
int cpb_vpsi_kpt(cpb_plan* plan, const void* c0, void* c2, long ld, int nstate, const double* f,
                 const double* hgkp, const double* hgkm, const double* vpot, int ngroups, int my_group,
                 unsigned flags);

/* ---- Hartree-Fock exchange ----------------------------------------------------------
 *   cpb_hfx_dev   SUBROUTINE hfx_old(c0,c2,f,psia,nstate,ehfx,vhfx) (hfx_utils.mod.F90:80-965) for func1%mhfx = 1
 *                 at the Gamma point without LSD and without Wannier / integral screening (hfxc3%twscr =
 *                 .FALSE.), one task, one group: every occupied state (f >= 1e-6) with itself (hfxaa :1203-1260)
 *                 and with every other occupied state (hfxab / hfxab2 :1034-1201): pair density psi_a psi_b /
 *                 omega -> dense forward transform on the PAIR-DENSITY set -> vpotg = -pf scgx rho(G), the pair
 *                 energy -> dense inverse -> v(r) (psi_a + i psi_b) -> sparse forward transform on the
 *                 wavefunction set -> c2a, c2b updated.  Two plans on the same mesh and device: `plan` built
 *                 from the wavefunction sphere (nzfs / inzs), `plan_dens` from the vectors of the pair-density
 *                 set (nzff / inzf, jhg = its ngw); scgx_dev: the Coulomb kernel of that set (cppt scgx), jhg
 *                 doubles.  pfl: 0.25 (times func3%phfx for a hybrid functional).  On return c2 += C2_hfx,
 *                 *ehfx = the exchange energy (:905), *vhfx = sum_i dotp(c0_i, c2_i) of the updated c2 (:907-909).
 *                 The real-space states stay in HBM for the duration of the call (8 nnr1 bytes per occupied
 *                 state, the reference's rswfx).  Not covered (the shim takes the original path): LSD, k-points,
 *                 screening (twscr / twfc), the 2-D task grid over pairs, hfxpsi / hfxrpa. */
int cpb_hfx_dev(cpb_plan* plan, cpb_plan* plan_dens, const void* c0_dev, void* c2_dev, long ld, int nstate,
                const double* f, const double* scgx_dev, double pfl, double* ehfx, double* vhfx, unsigned flags,
                void* stream);
/* host-pointer form (what the Fortran shim binds): all arrays in host memory, staged per call */
int cpb_hfx(cpb_plan* plan, cpb_plan* plan_dens, const void* c0, void* c2, long ld, int nstate, const double* f,
            const double* scgx, double pfl, double* ehfx, double* vhfx, unsigned flags);

/* ---- cross-group collectives over NVLink peer memory ----------------------------------------
 * One process per GPU (the reference's CP_GROUPS layout, one group per MPI rank).  Each rank
 * creates a SEGMENT of device memory, the ranks exchange the 64-byte handles with whatever they
 * have (MPI_Allgather over cp_inter_grp in the Fortran host; torch.distributed in the Python
 * harness) and map each other's segments (CUDA IPC).  Arrays that take part in a collective live
 * inside the segment (rhoe, vpot: pass cpb_peer_local_ptr() + offset to cpb_rhoofr_dev / cpb_vpsi_dev).
 *   cpb_peer_allreduce_f64  in-place sum over all ranks = cp_grp_redist(rhoe,nnr1,nlsd), i.e. mp_sum
 *                           over parai%cp_inter_grp (rhoofr_utils.mod.F90:457-461,
 *                           cp_grp_utils.mod.F90:98-120).  Deterministic (fixed rank order): all
 *                           ranks end with bit-identical data.
 *   cpb_peer_bcast_f64      rank `src`'s array to every rank (V(r) once per step).
 *   cpb_peer_allgather_f64  in place: block q (counts[q] doubles; the blocks lie back to back from
 *                           `offset` on) is valid on rank q on entry and on every rank on return.
 *   cpb_peer_redist_c2      cp_grp_redist(C2_vpsi) of a device-resident run (vpsi_utils.mod.F90:708-712,
 *                           forces_driver.mod.F90:283): the reference sums zero-padded full C2 arrays over
 *                           the groups, which is an all-gather of the owned state blocks.  The (ld, nstate)
 *                           COMPLEX*16 array starts `offset` doubles into the segment; the blocks are the
 *                           part_1d blocks of the ranks (part_1d.mod.F90:22-57).
 *   cpb_peer_allreduce_scalars  n <= 8 host doubles summed over the ranks in rank order (the group-partial
 *                           ekin / rsum_g / rsum_r of cpb_rhoofr_dev; the reference computes them redundantly
 *                           on every group, rhoofr_utils.mod.F90:178).  Synchronises the stream.
 * offset / n / counts are in doubles and even.  Calls are collective: every rank must make the same
 * sequence of calls.  cpb_peer_allreduce_f64 / cpb_peer_bcast_f64 / cpb_peer_allgather_f64 only ENQUEUE
 * their kernels on `stream` (later work on the same stream sees the result).  A rank that fails to show
 * up at a barrier within the timeout (cpb_peer_set_timeout_ms; default 20 s, or CPB_PEER_TIMEOUT_MS in
 * the environment) does not hang the device: the waiting ranks raise a sticky error word in EVERY
 * segment, every later collective kernel of the segment leaves its data untouched, and the segment is
 * unusable from then on.  The error is reported by the synchronising calls - cpb_peer_barrier,
 * cpb_peer_allreduce_scalars and cpb_peer_check return CPB_ERR_CUDA - so a caller MUST call
 * cpb_peer_check (or one of the other two) after the collectives of a step before it trusts the data.
 * cpb_peer_destroy unmaps the peers and frees the own segment: call it only after every rank has
 * passed a final cpb_peer_barrier (a peer may otherwise still be reading).  The mapped pointers of a
 * destroyed segment dangle: drop every view of the segment first. */
#define CPB_PEER_HANDLE_BYTES 64
typedef struct cpb_peer cpb_peer;
const char* cpb_peer_last_error(void);
int cpb_peer_create(cpb_peer** seg, int device, int rank, int world, size_t bytes, void* handle_out);
int cpb_peer_connect(cpb_peer* seg, const void* all_handles /* world * CPB_PEER_HANDLE_BYTES */);
void* cpb_peer_local_ptr(cpb_peer* seg);
int cpb_peer_barrier(cpb_peer* seg, void* stream);
int cpb_peer_check(cpb_peer* seg, void* stream);
int cpb_peer_allreduce_f64(cpb_peer* seg, size_t offset, size_t n, void* stream);
int cpb_peer_bcast_f64(cpb_peer* seg, size_t offset, size_t n, int src, void* stream);
int cpb_peer_allgather_f64(cpb_peer* seg, size_t offset, const size_t* counts /* world */, void* stream);
int cpb_peer_redist_c2(cpb_peer* seg, size_t offset, long ld, int nstate, void* stream);
int cpb_peer_allreduce_scalars(cpb_peer* seg, double* vals, int n, void* stream);
int cpb_peer_set_timeout_ms(cpb_peer* seg, double ms);
int cpb_peer_destroy(cpb_peer* seg);

/* ---- meta-GGA (cntl%ttau): kinetic-energy density and its potential ---------------------------
 *   cpb_tauofr_dev   SUBROUTINE tauofr(c0,psi,nstate) (tauofr_utils.mod.F90:42-111): tau(r) = sum_i
 *                    f_i tpiba2 / (2 omega) |grad psi_i(r)|^2 for the group's block, three sparse inverse
 *                    transforms per state pair (dpsisc :113-137 scales the coefficients by +-gk(k,ig);
 *                    tauadd :139-173).  tau_dev: (nnr1, nlsd), zeroed first (:80); with nsup >= 0 (LSD)
 *                    column 1 = alpha, column 2 = beta (tauadd's ispin1/ispin2).  cp_grp_redist(tau) stays
 *                    with the caller (:101-105; cpb_peer_allreduce_f64).
 *   cpb_vtaupsi_dev  SUBROUTINE vtaupsi(c0,c2,f,psi,nstate,ispin) (vtaupsi_utils.mod.F90:38-92): c2 -=
 *                    f tpiba2 / 4 * gk(k,ig) * unpack(FFT[vtau * d_k psi]) summed over k (taupot :94-129,
 *                    ftauadd :131-165); always accumulates into c2.  vtau_dev: (nnr1, ispin).
 * gk_dev: the cppt array gk(3,ngw) (column-major, Cartesian components of G in units of tpiba; gk(:,1) =
 * 0 when G=0 is the first vector).  nsup < 0: no LSD.  The pair that straddles the spin boundary runs as
 * two single states (same linear map as taupot's mixed branch). */
int cpb_tauofr_dev(cpb_plan* plan, const void* c0_dev, long ld, int nstate, const double* f, int nsup,
                   const double* gk_dev, int ngroups, int my_group, double* tau_dev, unsigned flags,
                   void* stream);
int cpb_vtaupsi_dev(cpb_plan* plan, const void* c0_dev, void* c2_dev, long ld, int nstate, const double* f,
                    int nsup, const double* gk_dev, const double* vtau_dev, int ngroups, int my_group,
                    unsigned flags, void* stream);
/* host-pointer forms (what the Fortran shim binds): all arrays in host memory, staged per call */
int cpb_tauofr(cpb_plan* plan, const void* c0, long ld, int nstate, const double* f, int nsup,
               const double* gk, int ngroups, int my_group, double* tau, unsigned flags);
int cpb_vtaupsi(cpb_plan* plan, const void* c0, void* c2, long ld, int nstate, const double* f, int nsup,
                const double* gk, const double* vtau, int ngroups, int my_group, unsigned flags);

/* number of kernel launches issued by this plan since creation (bench.py's gpu_launches) */
long cpb_plan_launch_count(const cpb_plan* plan);

/* Per-kernel-class device timing: when profiling is on, every launch is bracketed by CUDA events
 * on its stream; totals (ms) and launch counts are accumulated per class.  Used by bench.py for
 * the roofline of the dominant kernel; the counterpart of the reference's tiset/tihalt timers
 * around invfftn/fwfftn/vpsi/rhoofr (timer.mod.F90:39-233). */
enum {
  CPB_K_X_INV = 0, /* gather + pack + x inverse */
  CPB_K_Y_INV = 1,
  CPB_K_Z_RHO = 2, /* z inverse + density      */
  CPB_K_Z_VPSI = 3, /* z inverse * V * z forward */
  CPB_K_Y_FWD = 4,
  CPB_K_X_FWD = 5, /* x forward (scaled)       */
  CPB_K_KIN = 6,
  CPB_K_SUM = 7,
  CPB_K_UNPACK = 8, /* unpack + kinetic term + occupation + c2 update */
  CPB_K_DENSE = 9,  /* dense-transform extras: real-field z passes, G gather, ppener */
  CPB_NKINDS = 10
};
int cpb_plan_set_profiling(cpb_plan* plan, int on);
/* Batches of a call alternate between `n` work spaces/streams, 1 <= n <= the number allocated at
 * plan creation (env CPB_STREAMS, default 2: the tail of one batch's kernels runs beside the head of the next
 * batch's; 1 = all kernels serialised on one stream).  Results are bit-identical either way. */
int cpb_plan_set_streams(cpb_plan* plan, int n);
int cpb_plan_get_kernel_times(cpb_plan* plan, double* ms /*[CPB_NKINDS]*/, long* counts /*[CPB_NKINDS]*/,
                              int reset);

#ifdef __cplusplus
}
#endif
#endif /* CPB200_H */

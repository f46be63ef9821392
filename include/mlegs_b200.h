/*
 * mlegs_b200 -- C ABI of the B200-native MLegS hot path.
 *
 * The reference (UCBCFD/MLegS v1.1.3) has no FFI: its seam is the Fortran-2008
 * module/submodule split (src/modules/mlegs_scalar.f90 declares the interfaces,
 * src/submodules/mlegs_scalar_{init,dist,ops}.f90 hold the bodies; Makefile.dep:21-25).
 * A replacement submodule keeps mlegs_scalar.f90 unchanged and forwards each
 * `module procedure` to the entry point below that cites it (see INTEGRATION.md for
 * the ISO_C_BINDING interface block).  "ops" = src/submodules/mlegs_scalar_ops.f90,
 * "dist" = src/submodules/mlegs_scalar_dist.f90, "sinit" = src/submodules/mlegs_spectfm_init.f90.
 *
 * Conventions
 *  - every function returns 0 on success or an MLEGS_E_* code; mlegs_b200_last_error()
 *    then holds the reference's own `stop '...'` text for that failure.
 *  - a field is complex(p8) e(loc_sz(1),loc_sz(2),loc_sz(3)), column-major, resident in HBM;
 *    `e` is a device pointer owned by the library (mlegs_b200_field_alloc/free).
 *  - no torch / C++ types cross this boundary: plain pointers, ints, doubles.
 *  - there is no CPU fallback: without a CUDA device every compute entry fails with
 *    MLEGS_E_CUDA.
 */
#ifndef MLEGS_B200_H
#define MLEGS_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

enum {
  MLEGS_OK = 0,
  MLEGS_E_ARG = 1,      /* precondition failure -> the reference's `stop '<msg>'`        */
  MLEGS_E_STATE = 2,    /* library not initialised / wrong space                          */
  MLEGS_E_CUDA = 3,     /* CUDA runtime error (includes "no device")                      */
  MLEGS_E_COMM = 201,   /* reference error_flag_comm (modules/mlegs_envir.f90:40)         */
  MLEGS_E_MISC = 205    /* reference error_flag_misc: non-finite state (check_stability)  */
};

/* modules/mlegs_base.f90 globals that the hot path reads */
typedef struct mlegs_params {
  int nr, np, nz;
  int nrchop, npchop, nzchop;
  double ell, zlen;
  double visc;
  int hyperpow;
  double hypervisc;
  int is_svv;
  double svv_cutoff, svv_target, svv_strength, svv_relax;
} mlegs_params;

/* mirror of type(scalar), modules/mlegs_scalar.f90:15-49.  Updated in place. */
typedef struct mlegs_field {
  void *e;                 /* device pointer, complex(p8), column-major loc_sz              */
  int glb_sz[3], loc_sz[3], loc_st[3], axis_comm[3];
  double ln;
  int nrchop_offset, npchop_offset, nzchop_offset;
  char space[4];           /* "PPP","PFP","FFP","FFF" (+"PFF" inside vec2tp), NUL-terminated */
} mlegs_field;

/* ---- library / device ---------------------------------------------------------------- */
const char *mlegs_b200_last_error(void);
int mlegs_b200_version(void);
/* cudaStream_t the library launches on (NULL = legacy default stream). */
int mlegs_b200_set_stream(void *cuda_stream);
int mlegs_b200_device_sync(void);
/* number of kernels this library has launched since the last reset (bench.py's gpu_launches) */
long long mlegs_b200_launch_count(int reset);

/* per-kernel CUDA-event timing on the launching stream; report is a JSON object
 * {"kernel": {"launches": n, "ms": total}, ...} and clears the records */
int mlegs_b200_prof_enable(int on);
int mlegs_b200_prof_report(char *buf, size_t nbuf);
/* measured FP64 tensor-pipe (DMMA m8n8k4) throughput of this device in TFLOP/s: the roofline denominator of the
 * Legendre kernels (MEASURED_PEAKS.json has no FP64 entry) */
int mlegs_b200_dmma_peak(double *tflops);

/* ---- transform kit: host tables (tfm%init(), sinit:6-154) ---------------------------- */
/* sizes: x,w,ln,r: nr;  lognorm: (nrchop+14)*npchop;  pf: (nr/2)*(nrchop+14)*npchop;
 * at0,at1: nrchop;  ak: nz.  All column-major like the Fortran arrays.  Host code, runs the
 * three-term recurrence in binary128 where the reference uses 50-digit FM.                 */
int mlegs_b200_tfm_tables(const mlegs_params *p, double *x, double *w, double *ln, double *r,
                          double *lognorm, double *pf, double *at0, double *at1, double *ak);

/* The same through an on-disk cache (sinit:254-300 is the reference's documented slow phase):
 * <cache_dir>/mlegs_tables_<nr>_<nrchop>_<npchop>.bin holds x, w, lognorm, pf, at0, at1 (they depend on that triple
 * only) with a checksum; a missing or damaged file is rebuilt and rewritten atomically.  cache_dir NULL or "" builds
 * without caching.  *from_cache (may be NULL) reports whether the file was used. */
int mlegs_b200_tfm_tables_cached(const mlegs_params *p, const char *cache_dir, double *x, double *w, double *ln,
                                 double *r, double *lognorm, double *pf, double *at0, double *at1, double *ak,
                                 int *from_cache);

/* Upload the kit once; tables stay resident in HBM (replaces the host-side use of the global
 * `tfm` by ops:*).  rank/nranks describe the slab decomposition (dist:508-578 with
 * dims = (/nranks,1/)); peer buffers are attached later by mlegs_b200_dist_attach.       */
int mlegs_b200_init(const mlegs_params *p, const double *x, const double *w,
                    const double *lognorm, const double *pf, const double *at0,
                    const double *at1, int rank, int nranks);
int mlegs_b200_finalize(void);
int mlegs_b200_update_params(const mlegs_params *p);   /* visc, hypervisc, svv_* only */

/* ---- scalar storage: scalar_init / dealloc / copy (mlegs_scalar_init.f90:6-140) ------ */
/* space3 selects the layout: "PPP" -> axis_comm (1,0,2) (r sharded), anything else ->
 * (2,1,0) (m sharded); on one rank both are the full (nrdim,npdim,nzdim) block.           */
int mlegs_b200_field_alloc(mlegs_field *f, const char *space3);
/* on != 0: later field_alloc calls use cudaMallocManaged (preferred location = the device) so a Fortran
 * host can keep dereferencing s%e (apps/vortical_flow_3d.f90:136-137, 379); default is cudaMalloc. */
int mlegs_b200_use_managed(int on);
int mlegs_b200_field_free(mlegs_field *f);
int mlegs_b200_field_copy(mlegs_field *dst, const mlegs_field *src);      /* assignment(=) */
int mlegs_b200_field_zero(mlegs_field *f);
int mlegs_b200_field_upload(mlegs_field *f, const void *host_e);          /* whole local block */
int mlegs_b200_field_download(const mlegs_field *f, void *host_e);
int mlegs_b200_field_chop_offset(mlegs_field *f, int iof1, int iof2, int iof3);
/* pinned host staging (cudaHostRegister) for the host-buffer entry points below */
int mlegs_b200_host_register(void *host_ptr, size_t bytes);
int mlegs_b200_host_unregister(void *host_ptr);

/* ---- spectral transforms ------------------------------------------------------------- */
int mlegs_b200_trans(mlegs_field *s, const char to[3]);                   /* ops:157-235 */
/* trans() of n scalars that agree in space and chopping offsets -- the components of a vector field
 * (ops:1503-1505 transforms vr, vp, vz back to back) or the scalars of a multi-field app: every stage is ONE
 * launch over all of them (the scalar index is a grid dimension), results are bit-identical to n trans() calls.
 * A launch carries up to 32 scalars (larger n goes in groups); on several ranks a group is what one exchange epoch
 * carries (up to 8 slabs per window) and shares one fused exchange and one barrier per one-way transform.  Scalars in
 * mixed states fall back to a loop of trans(). */
int mlegs_b200_trans_many(int n, mlegs_field *const *s, const char to[3]);
/* Reference-facing call on a HOST array (the Fortran s%e): H2D, trans, D2H. */
int mlegs_b200_trans_host(void *host_e, const char from[3], const char to[3], double ln);
/* The same for n independent host arrays (a caller that transforms several scalars back to back, e.g. the three
 * components in ops:1503-1505): copies and transforms are pipelined over PCIe in both directions.  ln may be NULL. */
int mlegs_b200_trans_host_batch(int n, void *const *host_e, const char from[3], const char to[3], const double *ln);
/* scalar_exchange, dist:6-67 (slab layout: only (2,1)/(1,2) move data) */
int mlegs_b200_exchange(mlegs_field *s, int axis_old, int axis_new);

/* ---- masks / filters / far field ----------------------------------------------------- */
int mlegs_b200_chop(mlegs_field *s);                                      /* ops:6-41    */
int mlegs_b200_dealias(mlegs_field *s);                                   /* ops:43-70   */
int mlegs_b200_svv_filter(mlegs_field *s, double *gain);                  /* ops:72-155  */
int mlegs_b200_calcat0(const mlegs_field *s, double *out_nz_complex);     /* ops:237-272 */
int mlegs_b200_calcat1(const mlegs_field *s, double *out_nz_complex);     /* ops:274-309 */
int mlegs_b200_zeroat1(mlegs_field *s);                                   /* ops:311-325 */
/* smooth the far-field values: delsqp, radial synthesis, five smoothing passes over the radial tail of every retained
 * (m,k) line, radial analysis, idelsqp, zeroat1 */
int mlegs_b200_fftreat(mlegs_field *s);                                   /* ops:1002-1063 */

/* ---- spectral differential operators and solves -------------------------------------- */
int mlegs_b200_delsqp(mlegs_field *s);                                    /* ops:327-366 */
int mlegs_b200_idelsqp(mlegs_field *s);                                   /* ops:368-416 */
int mlegs_b200_xxdx(mlegs_field *s);                                      /* ops:418-463 */
int mlegs_b200_del2h(mlegs_field *s);                                     /* ops:465-518 */
int mlegs_b200_del2(mlegs_field *s);                                      /* ops:520-573 */
int mlegs_b200_idel2(mlegs_field *s, int have_preln, double preln);       /* ops:575-760 */
int mlegs_b200_ihelm(mlegs_field *s, double alpha);                       /* ops:791-854 */
/* (del^2 + alpha) s, ops:762-789.  The reference builds the result in a local scalar and drops it (no write-back), so
 * its helm is a no-op on s; this entry returns the documented operator (the inverse of ihelm). */
int mlegs_b200_helm(mlegs_field *s, double alpha);
int mlegs_b200_helmp(mlegs_field *s, int power, double alpha, double beta);   /* ops:856-903 */
int mlegs_b200_ihelmp(mlegs_field *s, int power, double alpha, double beta);  /* ops:905-1000 */

/* The LU factors of the (m,k) band systems depend on the operator only (power, alpha, beta, truncation), not on
 * the right-hand side: by default they are kept in HBM after the first solve and later solves run the two
 * substitutions only (bit-identical results; the reference re-factors in every call, ops:953-984).  on = 0
 * restores factor-per-call and frees the cache. */
int mlegs_b200_solve_cache(int on);

/* ---- time integrators ---------------------------------------------------------------- */
int mlegs_b200_fefe(mlegs_field *s, const mlegs_field *nl, double dt);    /* ops:1065-1094 */
int mlegs_b200_febe(mlegs_field *s, const mlegs_field *nl, double dt);    /* ops:1157-1198 */
int mlegs_b200_abcn(mlegs_field *s, mlegs_field *s_p, mlegs_field *nl,
                    mlegs_field *nl_p, double dt);                        /* ops:1200-1262 */

/* ops:1096-1155; is_2nd_svis_p != 0: the second argument already is the previous viscous term.  Quirk kept: in the
 * inviscid branch the reference zeroes svis instead of svis_p (ops:1135). */
int mlegs_b200_abab(mlegs_field *s, mlegs_field *s_p, mlegs_field *nl, mlegs_field *nl_p, double dt,
                    int is_2nd_svis_p);

/* ---- vector-field operations --------------------------------------------------------- */
int mlegs_b200_vecprod(mlegs_field *vr, mlegs_field *vp, mlegs_field *vz,
                       const mlegs_field *ur, const mlegs_field *up,
                       const mlegs_field *uz);                            /* ops:1264-1306 */
int mlegs_b200_vec2tp(const mlegs_field *vr, const mlegs_field *vp, const mlegs_field *vz,
                      mlegs_field *psi, mlegs_field *chi);                /* ops:1308-1453 */
int mlegs_b200_tp2vec(const mlegs_field *psi, const mlegs_field *chi,
                      mlegs_field *vr, mlegs_field *vp, mlegs_field *vz); /* ops:1455-1545 */
int mlegs_b200_tp2curlvec(const mlegs_field *psi, const mlegs_field *chi,
                          mlegs_field *wr, mlegs_field *wp, mlegs_field *wz); /* ops:1547-1560 */

/* ---- whole-array helpers the apps do with Fortran array syntax on s%e ---------------- */
/* y%e = a*x%e + b*y%e (e.g. apps/vortical_flow_3d.f90:136-137, 379) */
int mlegs_b200_axpby(mlegs_field *y, double a, const mlegs_field *x, double b);
/* all(ieee_is_finite(s%e)) -- check_stability, apps/vortical_flow_3d.f90:397-409 */
int mlegs_b200_is_finite(const mlegs_field *s, int *all_finite);

/* ---- on-device initial conditions / diagnostics of the vortical-flow apps ----------- */
/* The reference fills a GLOBAL array on every rank and scatters it (`disassemble`); here each rank fills its own
 * slab in HBM.  s must be in PPP.
 * s%e(i,j,k) = sum_c ((-exp(-d_c^2)) * mul / div) / (1 - x_i)^2 on both packed azimuthal lanes, d_c the distance of
 * the collocation point from centre (xo[c], yo[c]); rows >= nr, columns >= np/2 are zero
 * (apps/vortical_flow_3d.f90:279-294, 305-318).  ran_noise != 0 adds uniform(-1,1)*ran_noise inside r < ell of the
 * last centre from a counter-based generator (the reference's rand() is clock-seeded, :467-478). */
int mlegs_b200_gauss_vortices(mlegs_field *s, int ncentres, const double *xo, const double *yo, double mul, double div,
                              double ran_noise, unsigned long long seed);
/* uniform_z_fld, apps/vortical_flow_3d.f90:328-351: cmplx(re, im) on the physical points, zero on the padding */
int mlegs_b200_fill_physical(mlegs_field *s, double re, double im);
/* qvort_dist_tp, apps/vortical_flow_3d.f90:258-326: psi, chi (FFF in, FFF out) of two q-vortices at x = -2, +2 */
int mlegs_b200_qvort_dist_tp(mlegs_field *psi, mlegs_field *chi, double q, double ran_noise, unsigned long long seed);
/* the field of save_vort_mag, apps/vortical_flow_3d.f90:411-447: (wr,wp,wz) = tp2curlvec(psi,chi), vormag = |w| per
 * azimuthal lane; wr, wp, wz, vormag are PPP scalars supplied by the caller */
int mlegs_b200_vort_mag(const mlegs_field *psi, const mlegs_field *chi, mlegs_field *wr, mlegs_field *wp,
                        mlegs_field *wz, mlegs_field *vormag);

/* ---- field I/O in the reference's file formats ---------------------------------------- */
/* msave_scalar / mload_scalar, submodules/mlegs_scalar_io.f90:6-250.  is_global != 0: ONE file holding the global
 * array (binary stream: int32 n1,n2,n3 | complex(p8) a | real(p8) ln | int32 x3 chop offsets | character(3) space;
 * formatted: 1PE24.15E3 records).  The reference assembles the array on rank 0 (dist:70-203) and writes there; here
 * rank 0 lays the file out and every rank pwrite()s / pread()s its own slab at its byte offsets, so no rank ever
 * holds the global array.  is_global == 0: one file per rank, `<fn>_<rank>`, with glb_sz/loc_sz/loc_st in front. */
int mlegs_b200_msave(const mlegs_field *s, const char *fn, int is_binary, int is_global);
int mlegs_b200_mload(const char *fn, mlegs_field *s, int is_binary, int is_global);
/* The host halves of the two calls above (no CUDA): write / read the local block `host_e` described by `meta`
 * (glb_sz, loc_sz, loc_st, ln, offsets, space).  create != 0 creates the file with header and trailer first (rank 0
 * of a global file; every rank of a per-rank file); the caller orders create before the other ranks' writes. */
int mlegs_b200_msave_part(const mlegs_field *meta, const void *host_e, const char *fn, int is_binary, int is_global,
                          int rank, int create);
int mlegs_b200_mload_part(const char *fn, mlegs_field *meta, void *host_e, int is_binary, int is_global, int rank);

/* ---- multi-GPU (one process per GPU, slab over m) ------------------------------------ */
/* Ownership of the azimuthal wavenumbers on several ranks.  The reference's decompose() (mlegs_envir_mpi.f90:6-31) hands
 * out contiguous blocks; the work per column falls linearly with m (nn(m) = nrchop - m), so blocks leave rank 0 with
 * 1.4x the mean radial-transform / axial-FFT / solve work at 8 ranks.  Here rank q owns m = q, q + P, q + 2P, ...:
 * an m-distributed block has loc_st(2) = q (its first column) and its local column j holds global m = loc_st(2) +
 * stride * j, stride = P.  Blocks that hold every column (PPP layout, one rank) have stride 1.  Everything inside the
 * library (chop, transforms, operators, exchanges, msave/mload) follows this map; a host that indexes s%e by m itself
 * must use it too. */
int mlegs_b200_dist_m_stride(const mlegs_field *s, int *stride);
/* CUDA-IPC plumbing for the fused FFT+transpose kernels: each rank exports the handle of its
 * exchange window, the host side all-gathers the 64-byte handles (torch.distributed / MPI)
 * and attaches them.  Replaces MPI_Alltoallw + derived datatypes (dist:468-504).            */
int mlegs_b200_dist_window(void **dev_ptr, size_t *bytes, unsigned char handle64[64]);
int mlegs_b200_dist_attach(const unsigned char *handles64_all_ranks);
int mlegs_b200_dist_detach(void);
/* Host-only exchange plan (no CUDA needed): destination rank and linear index in that rank's new local block
 * for every element of rank `rank`'s local block in memory order; dir 0 = exchange(2,1), 1 = exchange(1,2) (columns
 * at their global positions: the user-visible s%exchange), 2 = exchange(1,2) into the transit layout of the exchanges
 * fused into a transform (columns grouped by owning rank; the azimuthal FFT that follows reads them permuted).
 * The device put kernels run the same addressing code (dist:395-504's subarray datatypes). */
int mlegs_b200_dist_put_map(int dir, int rank, int nranks, int nrdim, int npdim, int nz, int *dst_rank,
                            long long *dst_index);
/* Host-only plan of the STAGED exchange(1,2) (no CUDA needed): for every element of rank `rank`'s local
 * (nrdim, m_cnt, nz) block in memory order, its index inside the local staging buffer the Legendre synthesis writes
 * (stage_index), and -- for every staging index -- the rank and linear index the ship kernel moves it to
 * (ship_rank, ship_index; both of length nrdim*m_cnt*nz).  Composing the two must equal dist_put_map(dir = 2). */
int mlegs_b200_dist_stage_map(int rank, int nranks, int nrdim, int npdim, int nz, long long *stage_index,
                              int *ship_rank, long long *ship_index);
/* Sum n HOST doubles over all ranks on the library's own peer windows (the app-level MPI_Allreduce of
 * check_stability, apps/vortical_flow_3d.f90:404); identical result on every rank; no-op on one rank. */
int mlegs_b200_dist_allreduce(double *host_inout, int n);

#ifdef __cplusplus
}
#endif
#endif /* MLEGS_B200_H */

/* gfb.h -- C ABI of libgfb, the B200-native (sm_100a) pseudospectral timestep library.
 *
 * This is the drop-in boundary for the GeophysicalFlows.jl / FourierFlows.jl hot path
 * (SURVEY.md section 8b).  Plain C: raw device pointers, sizes, a CUDA stream; no torch types.
 * Every entry point returns 0 on success and a negative gfb_status on error (never aborts);
 * gfb_last_error() returns the message.  All work is enqueued asynchronously on `stream`
 * (a cudaStream_t passed as void*; NULL = default stream); only entry points that return
 * host scalars (gfb_diag) synchronise that stream.  A context is not thread-safe.
 *
 * Array conventions (identical bytes to the Julia arrays of the reference):
 *   spectral  ComplexF64[nkr, nl(, nlayers)]  ==  double2 a[layer][l][kr]   dense, nkr = nx/2+1
 *   physical  Float64[nx, ny(, nlayers)]      ==  double  a[layer][y][x]    dense
 *   wavenumbers as FourierFlows' TwoDGrid: kr = 2*pi/Lx * (0..nx/2), l = 2*pi/Ly * fftfreq(ny)
 *
 * Reference interfaces replaced (file:line under /root/reference, or [FF] = FourierFlows.jl,
 * not vendored, cited at the call site that pins it):
 *   gfb_create/gfb_destroy   Problem(dev; ...)          src/twodnavierstokes.jl:71-100, src/surfaceqg.jl:65-93,
 *                                                       src/singlelayerqg.jl:80-121, src/multilayerqg.jl:87-135
 *                            [FF] TimeStepper ctor      (tables: L, filter, ETD coefficients)
 *   gfb_rfft / gfb_irfft     mul!/ldiv!(.., grid.rfftplan, ..)   src/twodnavierstokes.jl:258-270
 *                            fwdtransform!/invtransform!         src/multilayerqg.jl:471,478
 *   gfb_dealias              [FF] dealias!(sol, grid)            src/twodnavierstokes.jl:287
 *   gfb_calcN                calcN!(N, sol, t, clock, vars, params, grid)
 *                                                       src/twodnavierstokes.jl:286-294, src/surfaceqg.jl:289-297,
 *                                                       src/singlelayerqg.jl:443-451, src/multilayerqg.jl:651-663
 *   gfb_step / gfb_stage     [FF] stepforward!(prob, nsteps)     test/test_twodnavierstokes.jl:251
 *   gfb_diag                 [FF] parsevalsum/parsevalsum2 + energy/enstrophy/... src/twodnavierstokes.jl:374-393,
 *                                                       src/surfaceqg.jl:390-416, src/singlelayerqg.jl:548-599,
 *                                                       src/multilayerqg.jl:918-951
 *   gfb_streamfunction_from_pv / gfb_pv_from_streamfunction
 *                            pv_streamfunction_kernel!           src/multilayerqg.jl:495-583
 *   gfb_updatevars           updatevars!(prob)                   src/twodnavierstokes.jl:331-345 (and analogues)
 */
#ifndef GFB_H
#define GFB_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct gfb_ctx gfb_ctx;

enum gfb_status {
  GFB_OK = 0,
  GFB_ERR_INVALID = -1,     /* bad argument / unsupported configuration */
  GFB_ERR_CUDA = -2,        /* CUDA runtime error (message has the CUDA string) */
  GFB_ERR_ALLOC = -3,
  GFB_ERR_UNSUPPORTED = -4  /* e.g. nx that is not a power of two in [32, 16384] */
};

/* which calcN! family */
enum gfb_kind {
  GFB_KIND_SINGLE = 0, /* TwoDNavierStokes, SurfaceQG, SingleLayerQG with constant U:  psi_h = -G * sol */
  GFB_KIND_LAYERED = 1, /* MultiLayerQG (any nlayers) and SingleLayerQG with U(y):       psi_h = S^-1 sol,
                          extra (u+U) dQ/dx + v dQ/dy term, optional bottom drag */
  GFB_KIND_QL = 2      /* BarotropicQGQL (src/barotropicqgql.jl:293-333): psi_h = -G * sol as GFB_KIND_SINGLE; calcN! keeps
                          the eddy-mean products at kr != 0 and the eddy-eddy flux -d/dy <v'zeta'> at kr = 0 (zonal mean =
                          the kr = 0 column).  Config fields as GFB_KIND_SINGLE (eta and Nconst are not used by the reference's
                          calcN_advection! and must be NULL). */
};

enum gfb_stepper {
  GFB_FORWARD_EULER = 0,
  GFB_FILTERED_FORWARD_EULER = 1,
  GFB_RK4 = 2,
  GFB_FILTERED_RK4 = 3,
  GFB_ETDRK4 = 4,
  GFB_FILTERED_ETDRK4 = 5,
  GFB_AB3 = 6,
  GFB_FILTERED_AB3 = 7,
  GFB_LSRK54 = 8,          /* [FF] LSRK54TimeStepper: Carpenter-Kennedy five-stage fourth-order 2N-storage Runge-Kutta */
  GFB_FILTERED_LSRK54 = 9  /* the same with timestepper.filter applied at the end of the step */
};

/* All pointers in the config are HOST pointers, read (copied to the device) during gfb_create. */
typedef struct gfb_config {
  int struct_size;        /* = sizeof(gfb_config) of the header the CALLER was built against; gfb_create rejects a
                             mismatch (a binding whose struct mirror has drifted) instead of reading past the struct */
  int kind;               /* gfb_kind */
  int nx, ny, nlayers;    /* nx: power of two in [32, 16384]; ny: power of two in [32, 16384], or (one GPU) any even number in
                             [8, 1024] -- those columns run as direct transforms (csrc/cols_dft.cu; the reference's tests use 126, 66, 34) */
  double Lx, Ly;
  /* dealias! index ranges, 0-based half-open ([FF] getaliasedwavenumbers): columns [kr_alias0, nkr) and
   * rows [l_alias0, l_alias1) are zeroed.  kr_alias0 = nkr and l_alias0 = l_alias1 disables. */
  int kr_alias0, l_alias0, l_alias1;
  int stepper;            /* gfb_stepper */
  double dt;
  int L_is_complex;       /* 0: L and the ETD tables are double[...]; 1: double2[...] */
  const void* L;          /* equation.L        [nlayers][nl][nkr] */
  const double* filter;   /* timestepper.filter [nl][nkr] (same for all layers) or NULL */
  /* ETDRK4 tables ([FF] ETDRK4TimeStepper fields), same type/shape as L; NULL unless an ETDRK4 stepper */
  const void* expLdt;
  const void* exphalfLdt;
  const void* zeta;
  const void* alpha;
  const void* beta;
  const void* Gamma;
  /* GFB_KIND_SINGLE */
  const double* G;        /* [nl][nkr]: psi_h = -G*sol  (grid.invKrsq | params.psihfrombh | 1/(Krsq+1/ell^2)) */
  const double* eta;      /* [ny][nx] topographic PV added to q in the products, or NULL */
  const void* Nconst;     /* double2 [nlayers][nl][nkr] constant added to N (e.g. -i kr U eta_h), or NULL */
  /* GFB_KIND_LAYERED */
  const double* Sinv;     /* [nl][nkr][nlayers][nlayers] row-major: psi_h = Sinv * q_h   (params.S^-1) */
  const double* S;        /* same layout: q_h = S * psi_h (params.S); may be NULL if pv_from_streamfunction unused */
  const double* U;        /* [nlayers][ny] imposed zonal flow, or NULL */
  const double* Qx;       /* [nlayers][ny][nx] */
  const double* Qy;       /* [nlayers][ny][nx] */
  double mu;              /* bottom-layer linear drag (MultiLayerQG calcN!), 0 if none */
  int no_dealias_in_calcN; /* 1: MultiLayerQG.calcNlinear! (src/multilayerqg.jl:675-683, 744-779): calcN! never dealiases and drops the self-advection (u q, v q); GFB_KIND_LAYERED only */
  /* Multi-GPU slab decomposition (the reference is single-device).  nranks = 0 or 1: one GPU.
   * nranks = P > 1 (power of two, ny % (8P) == 0): this context owns
   *   - spectral side: column PAIRS dealt round-robin -- local pair p is global pair p*P + rank; every
   *     spectral table above and every spectral array passed later is the LOCAL dense slice [..][nl][nkl];
   *   - physical side: rows [rank*ny/P, (rank+1)*ny/P); eta/U/Qx/Qy are the local row slabs. */
  int rank, nranks;
} gfb_config;

const char* gfb_version(void);
int gfb_config_size(void); /* sizeof(gfb_config) of the LIBRARY: every binding asserts its own mirror against it */
const char* gfb_last_error(const gfb_ctx* ctx); /* ctx may be NULL: error of the last failed gfb_create */

int gfb_create(gfb_ctx** out, const gfb_config* cfg);
int gfb_destroy(gfb_ctx* ctx);

/* clock (FourierFlows Clock: t accumulates by repeated t += dt) */
int gfb_get_clock(const gfb_ctx* ctx, double* t, long long* step);
int gfb_set_clock(gfb_ctx* ctx, double t, long long step);

/* 2-D transforms over `batch` layers.  out: double2[batch][nl][nkr]; in: double[batch][ny][nx].
 * gfb_rfft is unnormalised, gfb_irfft scales by 1/(nx*ny) and does NOT destroy its input. */
int gfb_rfft(gfb_ctx* ctx, void* out_spec, const double* in_phys, int batch, void* stream);
int gfb_irfft(gfb_ctx* ctx, double* out_phys, const void* in_spec, int batch, void* stream);

/* dealias!(sol, grid): in place */
int gfb_dealias(gfb_ctx* ctx, void* sol, void* stream);

/* calcN!: dealiases `sol` IN PLACE, then N = nonlinear term (+ Nconst + drag) + Fh (Fh may be NULL). */
int gfb_calcN(gfb_ctx* ctx, void* N, void* sol, const void* Fh, void* stream);

/* stepforward!(prob, nsteps) with forcing frozen over each step (exact for stochastic=true and for
 * time-independent forcing).  prevsol (may be NULL) receives the dealiased pre-step state of the LAST
 * step taken, as addforcing! does (src/twodnavierstokes.jl:312-315). */
int gfb_step(gfb_ctx* ctx, void* sol, int nsteps, const void* Fh, void* prevsol, void* stream);

/* One stage of the current step (0-based); the host calls calcF! between stages for time-dependent
 * deterministic forcing.  gfb_nstages() stages make one step; the clock advances after the last one. */
int gfb_nstages(const gfb_ctx* ctx);
int gfb_stage(gfb_ctx* ctx, void* sol, int istage, const void* Fh, void* prevsol, void* stream);

/* weighted Parseval sums; see enum gfb_diag_kind.  Returns raw sums
 *   S = sum_l ( f[0,l] + f[nkr-1,l] + 2 sum_{0<k<nkr-1} f[k,l] )
 * (parsevalsum WITHOUT the Lx*Ly/(nx*ny)^2 factor) into out_host[0..nout). */
enum gfb_diag_kind {
  GFB_DIAG_ABS2 = 0,      /* |s|^2 */
  GFB_DIAG_INVK_ABS2 = 1, /* invKrsq |s|^2 */
  GFB_DIAG_K_ABS2 = 2,    /* Krsq |s|^2 */
  GFB_DIAG_KP_ABS2 = 3,   /* Krsq^p |s|^2, [0,0] excluded */
  GFB_DIAG_G_ABS2 = 4,    /* G |s|^2 */
  GFB_DIAG_G2_ABS2 = 5,   /* G^2 |s|^2 = |psi_h|^2 */
  GFB_DIAG_KG2_ABS2 = 6,  /* Krsq G^2 |s|^2 */
  GFB_DIAG_ABS2_PLUS = 7, /* |s + aux|^2 */
  GFB_DIAG_WORK = 8,      /* Re(s' conj(aux)), s' = (s + aux2)/2 if aux2 else s */
  GFB_DIAG_WORK_INVK = 9, /* invKrsq Re(s' conj(aux)) */
  GFB_DIAG_LAYERED_ENERGIES = 10 /* nlayers values Krsq|psi_j|^2 then nlayers-1 values |psi_j - psi_j+1|^2 */
};
int gfb_diag(gfb_ctx* ctx, int kind, const void* sol, const void* aux, const void* aux2, double p,
             double* out_host, int nout, void* stream);

/* psi_h = S^-1 q_h and q_h = S psi_h (GFB_KIND_LAYERED), or psi_h = -G q_h (GFB_KIND_SINGLE; the
 * inverse direction is then unsupported). */
int gfb_streamfunction_from_pv(gfb_ctx* ctx, void* psih, const void* qh, void* stream);
int gfb_pv_from_streamfunction(gfb_ctx* ctx, void* qh, const void* psih, void* stream);

/* updatevars!: dealias sol in place, then any non-NULL output gets the physical field
 * (q, psi, u, v: double[nlayers][ny][nx]; u WITHOUT the imposed U, as vars.u in updatevars!). */
int gfb_updatevars(gfb_ctx* ctx, void* sol, double* q, double* psi, double* u, double* v, void* stream);

/* ---- sharded stepping (nranks > 1; also valid on one GPU) -------------------------------------------------
 * One stage = gfb_stage_pre (spectral pre-pass + inverse y-FFT of the local columns -> wsend)
 *           -> all-to-all (caller: NCCL) -> gfb_stage_mid (row pass on the local rows: wrecv -> psend)
 *           -> all-to-all -> gfb_stage_post (forward y-FFT + stage update of the local columns).
 * wsend/wrecv (psend/precv) are caller-owned buffers of info[5] (info[6]) complex elements whose P equal
 * contiguous chunks are exactly the all-to-all send/receive blocks (8-row tiles, no packing).  The state lives
 * in the context between gfb_state_import and gfb_state_export (local dense [nlayers][nl][nkl]).
 * gfb_shard_info: info = {rank, nranks, nkl, padded local columns, local rows, W elements, P elements, pairs/rank}. */
int gfb_shard_info(const gfb_ctx* ctx, long long* info8);
int gfb_state_import(gfb_ctx* ctx, const void* sol_local, void* stream);
int gfb_state_export(gfb_ctx* ctx, void* sol_local, void* stream);
int gfb_stage_pre(gfb_ctx* ctx, int istage, void* wsend, void* stream);
int gfb_stage_mid(gfb_ctx* ctx, const void* wrecv, void* psend, int prune, void* stream);
int gfb_stage_post(gfb_ctx* ctx, int istage, const void* precv, const void* Fh_local, void* prevsol_local, int prune,
                   void* stream);

/* Lazy alias band for callers that step one call at a time (stochastic forcing: calcF! runs between the steps).
 * gfb_step / gfb_step_p2p always compute the aliased band of N -- which lands only in the aliased band of the
 * returned state and is wiped by the next dealias! -- in the LAST step of a call, so that everything a caller can
 * observe is exact.  While this switch is on the last step skips it too: the aliased band of the returned `sol` is
 * then unspecified.  Legal whenever the next thing that touches `sol` is dealias! (the reference's calcN! starts
 * with it, src/twodnavierstokes.jl:287); switch it off for the step after which diagnostics or output read `sol`. */
int gfb_set_defer_alias_band(gfb_ctx* ctx, int on);

/* "p2p" transport of a sharded context: instead of handing send buffers to an all-to-all, K1 and K2 store every
 * row / column pair straight into the buffer of the rank that consumes it (peer memory over NVLink), and a
 * flag barrier in peer memory separates the passes -- the exchange is fused into the kernels.
 *   gfb_comm_create   allocates this rank's receive buffer; handle64_out gets its 64-byte cudaIpcMemHandle_t
 *   gfb_comm_connect  opens the other ranks' buffers: `handles` = nranks x 64 bytes, entry d from rank d
 *   gfb_comm_connect_ptrs  same for contexts living in ONE process (raw pointers from gfb_comm_create)
 *   gfb_step_p2p      the sharded stepforward! (same arguments as gfb_step, local arrays); every rank must call it
 *   gfb_comm_status   synchronises and reports a barrier that gave up on a missing peer
 * Error model of the flag barriers (asynchronous, like CUDA's own): a barrier that gives up (GFB_BARRIER_SPINS polls,
 * default 4e7) writes its epoch into a mapped host word.  Every later barrier launch and the end of gfb_step_p2p read
 * that word on the host and return GFB_ERR_CUDA ("peer barrier timed out"); the failure is sticky for the context. */
int gfb_comm_create(gfb_ctx* ctx, void* handle64_out, void** local_ptr_out);
int gfb_comm_connect(gfb_ctx* ctx, const void* handles);
int gfb_comm_connect_ptrs(gfb_ctx* ctx, void* const* ptrs);
int gfb_comm_status(gfb_ctx* ctx);
int gfb_step_p2p(gfb_ctx* ctx, void* sol_local, int nsteps, const void* Fh_local, void* prevsol_local, void* stream);

/* Per-kernel device timing (CUDA events on the caller's stream around every launch), for the roofline
 * report of bench.py.  Classes: 0 = k_col_inv (K1), 1 = k_row (K2), 2 = k_col_fwd (K3), 3 = other.
 * gfb_profile_end synchronises the recorded events and returns summed milliseconds and launch counts. */
int gfb_profile_begin(gfb_ctx* ctx);
int gfb_profile_end(gfb_ctx* ctx, double* ms4, long long* launches4);

/* Which kernel variants the context dispatches to (for tests and the bench line):
 *   info[0] = 1: the row pass of the stepping path is TMA-staged (k_row_product_tma / k_row_product_ws); info[1] = launches
 *   info[2] = 1: the column pre-pass is the warp-specialised kernel (k_col_inv_ws); info[3] = its launches so far
 *   info[4] = 1: the row pass is the warp-specialised kernel (k_row_product_ws)
 *   info[5] = 1: the forward column pass + epilogue is the warp-specialised kernel (k_col_fwd_ws); info[6..7] reserved */
int gfb_path_info(const gfb_ctx* ctx, long long* info8);

/* bytes of device memory owned by the context */
long long gfb_device_bytes(const gfb_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* GFB_H */

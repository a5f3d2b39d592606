/* nremfc.h — C ABI of the B200-native Wilson-Cowan -> BOLD -> FC -> GoF hot path.
 *
 * Everything the reference does between "module attributes are set" and "one row of
 * goodness-of-fit numbers exists" (netwWilsonCowanPlastic.py:77-158, whole_sweep_both.py:78-94,
 * utils.py:42-50) is reachable through the entry points below.  The reference has no FFI of
 * its own (it is Python + numba); each entry point names the reference call it replaces.
 *
 * Conventions
 *   - every function returns 0 on success, a negative nrem_status otherwise; the message of
 *     the last failure on the calling thread is nrem_last_error().  No exceptions cross.
 *   - all array arguments are DEVICE pointers (cudaMalloc'd / torch CUDA tensors) unless the
 *     name starts with h_; outputs are caller-allocated; nothing is retained after return.
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).  Calls are
 *     asynchronous with respect to the host unless stated otherwise.
 *   - row-major ("C order") everywhere; shapes are written [outer, ..., inner].
 *   - there is NO CPU fallback: without a CUDA device every compute entry point fails with
 *     NREM_ERR_CUDA.
 */
#ifndef NREMFC_H
#define NREMFC_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NREM_ABI_VERSION 2

typedef enum nrem_status {
    NREM_OK = 0,
    NREM_ERR_ARG = -1,      /* bad shape / null pointer / unsupported size */
    NREM_ERR_CUDA = -2,     /* CUDA runtime error (no device, launch failure, ...) */
    NREM_ERR_UNSUPPORTED = -3
} nrem_status;

/* Model constants — the module attributes of netwWilsonCowanPlastic.py:23-57 that
 * wilsonCowan()/run() freeze at JIT time, plus the phase lengths len(timeTrans1),
 * len(timeTrans2), len(timeSim) (netwWilsonCowanPlastic.py:48-50). */
typedef struct nrem_wc_params {
    double a_ee, a_ie_0, a_ei, a_ii;   /* :23-25 */
    double tauE, tauI;                 /* :27 */
    double P, rhoE;                    /* :29,:32 */
    double rE, rI, mu, sigmaI;         /* :35-38 */
    double dtSim, sqdtD;               /* :46,:56 */
    double E0, I0;                     /* :90-91 */
    double tau_ip[3];                  /* :101,:111,:118 */
    int64_t n1, n2, n3;                /* Euler steps per phase */
    int32_t downsamp;                  /* int(dt/dtSim), :120 */
    int32_t nnodes;                    /* N = len(CM), :68 */
    uint64_t seed;                     /* key of the counter-based noise stream */
} nrem_wc_params;

int nrem_abi_version(void);
const char* nrem_last_error(void);
/* Number of visible CUDA devices (0 when there is none; never fails). */
int nrem_device_count(void);

/* ---- reference-shaped stages (float64) -------------------------------------------------- */

/* Replaces run() (netwWilsonCowanPlastic.py:86-137) for B independent simulations, float64.
 *   CM      [N,N]        structural connectivity
 *   G,sigmaE [B,N]       per-simulation, per-node coupling gain and excitatory slope
 *   streams [B]          replicate ids of the counter-based noise (ignored when noise != NULL)
 *   noise   NULL, or [B or 1, n1+n2+n3, N] values to add inside the sigmoid (already scaled by
 *           sqdtD — exactly what np.random.normal(0, sqdtD, N) returned at :80); noise_batch = 1
 *           shares one stream between all B simulations.
 *   Y       NULL or [B, nrec, 3, N]   state (E, I, a_ie) stored BEFORE step i when i % downsamp == 0
 *   final   NULL or [B, 3, N]         state after the last step                                  */
int nrem_wc_run_f64(const nrem_wc_params* p, const double* CM, const double* G, const double* sigmaE,
                    const uint64_t* streams, const double* noise, int noise_batch, int B, int64_t nrec,
                    double* Y, double* final_state, void* stream);

/* Same with per-node vectors for the scalar model parameters ("Any of them can be redefined as a vector of length nnodes",
 * netwWilsonCowanPlastic.py:21).  node_params: NULL or [NREM_NODE_PARAMS, N] (device) in the order
 * a_ee, a_ei, a_ii, tauE, tauI, P, rhoE, rE, rI, mu, sigmaI, a_ie_0; row k overrides the scalar of `p` for every node. */
#define NREM_NODE_PARAMS 12
int nrem_wc_run_f64_ex(const nrem_wc_params* p, const double* CM, const double* G, const double* sigmaE,
                       const double* node_params, const uint64_t* streams, const double* noise, int noise_batch, int B,
                       int64_t nrec, double* Y, double* final_state, void* stream);

/* Replaces wilsonCowan(t, X, sigmaE, mu, tau_ip, G) (netwWilsonCowanPlastic.py:77-83): one
 * derivative evaluation.  X [3,N], noise [N] (scaled), G/sigmaE [N] -> dX [3,N].            */
int nrem_wc_derivative_f64(const nrem_wc_params* p, const double* CM, const double* X, const double* G,
                           const double* sigmaE, const double* noise, double tau_ip, double* dX, void* stream);

/* Replaces BOLDModel.Sim(rE, nnodes, dt) (call site netwWilsonCowanPlastic.py:144).
 *   rE [B,T,N] -> bold [B,T,N]; explicit-Euler Balloon-Windkessel, bold[t] from the state before rE[t]. */
int nrem_bold_sim_f64(const double* rE, int B, int64_t T, int N, double dt, double* bold, void* stream);

/* Replaces the cut + filtfilt + decimate of simBOLD (netwWilsonCowanPlastic.py:145-156).
 *   bold [B,T,N]; the first Neq rows are dropped; zero-phase IIR (h_b, h_a: 5 coefficients each,
 *   host pointers, SciPy order, a[0] = 1) with SciPy's default odd padding of 15 and lfilter_zi
 *   start-up; every ds-th row is kept -> out [B, ceil((T-Neq)/ds), N].
 *   scratch: device, >= nrem_filt_scratch_bytes(B,T,N,Neq,ds) bytes.                           */
int64_t nrem_filt_scratch_bytes(int B, int64_t T, int N, int64_t Neq, int64_t ds);
int nrem_filtfilt_decimate_f64(const double* bold, int B, int64_t T, int N, int64_t Neq, int64_t ds,
                               const double* h_b, const double* h_a, double* out, void* scratch, void* stream);

/* Replaces np.corrcoef(BOLD.T) (whole_sweep_both.py:81).  bold [B,J,N] -> fc [B,N,N].  Any N: up to 128 nodes one CTA per
 * simulation with the matrix in shared memory, above that one CTA per 32 x 32 tile.                                   */
int nrem_fc_f64(const double* bold, int B, int64_t J, int N, double* fc, void* stream);

/* Replaces utils.get_all_metrics(sFC, empFC, data_range=1) for K targets (utils.py:42-50,
 * whole_sweep_both.py:83-86) and sFC.mean() (whole_sweep_both.py:94).
 *   fc [B,N,N], emp [K,N,N] -> gof [B,K,4] = (corr, euc, ssim, new_metric); meanfc [B] (may be NULL).
 * N >= 7 (the 7 x 7 SSIM window); above 128 nodes both matrices stay in global memory (one CTA per simulation and target). */
int nrem_gof_f64(const double* fc, const double* emp, int B, int K, int N, double data_range,
                 double* gof, double* meanfc, void* stream);

/* Replaces utils.kuramoto(BOLD) (utils.py:34-40, whole_sweep_both.py:93): Hilbert phases along time (scipy.signal.hilbert
 * semantics, length-J FFT), order parameter |mean_n exp(i phase)|, then its mean ("sync") and population std ("meta").
 *   bold [B,J,N] -> sync_meta [B,2];  scratch_g: device, >= J doubles;  2 <= J <= 1024.                        */
int nrem_kuramoto_f64(const double* bold, int B, int64_t J, int N, double* sync_meta, void* scratch_g, void* stream);

/* ---- fused sweep (the fast path; no counterpart in the reference) ------------------------- */

/* Per-simulation parameters of a G x sigma x seed x map sweep
 * (whole_sweep_both.py:68-72, whole_sweep_both_maps.py:104-108, run_many_seeds.py:115-119):
 *     G_i     = G0[b]     + dG[b]     * mapG[map_id[b]][i]
 *     sigma_i = sigma0[b] + dsigma[b] * mapS[map_id[b]][i]
 * Simulations are processed in tiles of NREM_TILE_SIMS; all simulations of one tile must share
 * map_id (the host layer pads).                                                               */
#define NREM_TILE_SIMS 128

typedef struct nrem_sweep_plan nrem_sweep_plan;   /* opaque; owns device scratch */

typedef struct nrem_sweep_opts {
    int32_t kernel;          /* 0 = auto, 1 = CUDA-core coupling, 2 = tcgen05 (TF32) coupling, 3 = tcgen05 3xTF32,            */
                             /* 5 / 6 = node-lane tcgen05 3xTF32 with 32 / 16 simulations per CTA (see nrem_sweep_kernel),      */
                             /* 7 = large-connectome integrator (one launch per Euler step, 16 <= nnodes <= 8192, one map pair): */
                             /*     what auto picks above 128 nodes                                                              */
    int32_t bold_f32;        /* 1 = Balloon-Windkessel state in float32 (default 0 = float64)      */
    int32_t chunk_samples;   /* stored samples per launch of the integrator (0 = default)           */
    int32_t want_fc;         /* 1 = also return the FC matrices                                      */
    int64_t Neq;             /* rows dropped before filtering (reference: 2000)                      */
    int64_t bold_downsamp;   /* decimation after filtering (reference: 1000)                         */
    double  bold_dt;         /* BOLD Euler step (reference: dt*downsamp = 0.04)                      */
    double  b[5], a[5];      /* band-pass coefficients, SciPy order                                  */
    int32_t welch_nperseg;   /* 0 = no spectrum; else segment length of signal.welch (reference: 4000): even, <= T,     */
                             /* nperseg/2 = 2^a 5^b <= 2560 and nperseg/2 a multiple of chunk_samples; 50 % overlap, Hann */
    int32_t reserved;
    double  welch_fs;        /* sampling rate of the stored samples, 1/dt (reference: 500 Hz)                 */
} nrem_sweep_opts;

/* 7 <= nnodes <= 8192.  Up to 128 nodes the register-resident integrators (kernels 1-6) run; above that (any other parcellation,
 * BASELINE configs[4]) the plan integrates with the large-connectome kernel of nrem_big_integrate_f32 and runs the same
 * BOLD -> filter -> FC -> GoF -> Kuramoto chain, FC / GoF in batches of simulations (n_maps must be 1).                    */
int nrem_sweep_create(const nrem_wc_params* p, const nrem_sweep_opts* o, int B, int n_maps, int K,
                      nrem_sweep_plan** plan);
int nrem_sweep_destroy(nrem_sweep_plan* plan);
int64_t nrem_sweep_device_bytes(const nrem_sweep_plan* plan);

/* Runs the whole pipeline for B simulations.
 *   CM [N,N] f64; mapG,mapS [n_maps,N] f64; G0,dG,sigma0,dsigma [B] f64; h_map_id [B] i32 (HOST pointer, may be NULL = all 0);
 *   streams [B] u64; emp [K,N,N] f64
 *   gof [B,K,4] f64; extra [B,4] f64 = (mean FC, sync, meta, peakfreq) — the last four columns of the reference's
 *   output row (whole_sweep_both.py:90-96); peakfreq is NaN unless opts.welch_nperseg > 0, sync/meta are NaN when the
 *   decimated BOLD has more than 1024 rows; fc NULL or [B,N,N] f64.
 * Equivalent to nrem_sweep_begin(homogeneous = -1) + nrem_sweep_advance(all) + nrem_sweep_finish: everything is enqueued on
 * `stream` and the call returns without waiting for the kernels, except for ONE stream synchronisation inside begin (the
 * homogeneity of the maps is read back to pick the kernel specialisation; pass the hint through nrem_sweep_begin to avoid it). */
int nrem_sweep_run(nrem_sweep_plan* plan, const double* CM, const double* mapG, const double* mapS,
                   const double* G0, const double* dG, const double* sigma0, const double* dsigma,
                   const int32_t* h_map_id, const uint64_t* streams, const double* emp,
                   double* gof, double* extra, double* fc, void* stream);

/* The same run cut into pieces (the time loop is sequential, so a long sweep can be advanced slice by slice with the state kept
 * in the plan; bench.py times such slices).  All three calls only enqueue work on `stream`.
 *   begin   : stages the inputs and rewinds the plan to Euler step 0.  homogeneous: 1 = every entry of mapG/mapS is exactly 1
 *             (scalar G and sigma per simulation: the specialised kernel), 0 = not, -1 = find out on the device (synchronises once).
 *   advance : enqueues at most max_chunks integrator launches (one launch = chunk_samples * downsamp Euler steps of every
 *             simulation, never crossing a phase boundary; a whole run has nrem_sweep_chunks_total of them) plus the
 *             BOLD/filter launches that consume their samples; *h_chunks_left (host, may be NULL) = launches still to go.
 *   finish  : backward filter pass, FC, GoF, sync/meta/peakfreq of a run whose integration is complete (outputs as nrem_sweep_run). */
/* Per-node vectors for the scalar model parameters of the NEXT runs of this plan ("Any of them can be redefined as a vector of
 * length nnodes", netwWilsonCowanPlastic.py:21): node_params = device [NREM_NODE_PARAMS, N] in the order of nrem_wc_run_f64_ex
 * (copied into the plan), or NULL to go back to the scalars of nrem_wc_params.  Selects the node-lane kernel.            */
int nrem_sweep_set_node_params(nrem_sweep_plan* plan, const double* node_params, void* stream);
/* The integrator kernel the plan resolved to (1 CUDA-core, 2 tcgen05 TF32, 3 tcgen05 3xTF32: 128 simulations per CTA, a
 * thread = one simulation x 24 nodes; 5 / 6 node-lane tcgen05 3xTF32: 32 / 16 simulations per CTA, a thread = one node x 8 / 4
 * simulations, nnodes <= 128).  opts.kernel = 0 picks 6 or 5 for batches too small to fill the SMs with 128-simulation tiles,
 * for nnodes > 96 and for per-node parameter tables, else 3; 7 = large-connectome integrator (nnodes > 128).                  */
int nrem_sweep_kernel(const nrem_sweep_plan* plan);
int nrem_sweep_begin(nrem_sweep_plan* plan, const double* CM, const double* mapG, const double* mapS,
                     const double* G0, const double* dG, const double* sigma0, const double* dsigma,
                     const int32_t* h_map_id, const uint64_t* streams, int homogeneous, void* stream);
int64_t nrem_sweep_chunks_total(const nrem_sweep_plan* plan);
int nrem_sweep_advance(nrem_sweep_plan* plan, int64_t max_chunks, int64_t* h_chunks_left, void* stream);
int nrem_sweep_finish(nrem_sweep_plan* plan, const double* emp, double* gof, double* extra, double* fc, void* stream);

/* Test hook for the deterministic stages of the sweep: instead of integrating, feed stored E samples (what run() records,
 * netwWilsonCowanPlastic.py:129-130) to the plan's own BOLD / forward-filter / spectrum kernels, then call nrem_sweep_finish.
 *   E [rows, N, Bpad] f32, simulation fastest, Bpad = B rounded up to NREM_TILE_SIMS; rows are consecutive stored samples
 *   continuing where the previous call stopped (after nrem_sweep_begin: row 0); every call but the last must bring a
 *   multiple of chunk_samples rows.                                                                                   */
int nrem_sweep_feed_samples(nrem_sweep_plan* plan, const float* E, int64_t rows, void* stream);

/* Optional device-side timing (CUDA events on the caller's stream and on tile group 0's stream).
 * nrem_sweep_get_profile blocks until the enqueued work has finished, fills h_out[4] (host) with
 * {device ms of the begin/advance/finish/run calls, integrator ms of tile group 0, its launches, number of tile groups}
 * accumulated since the previous get_profile (or set_profiling) call, and starts a new accumulation.             */
int nrem_sweep_set_profiling(nrem_sweep_plan* plan, int on);
int nrem_sweep_get_profile(nrem_sweep_plan* plan, double* h_out);

/* Test hooks for the sweep's integrator: advance B simulations n1+n2+n3 steps with kernel
 * variant `kernel` (codes of nrem_sweep_opts.kernel) and return the float32 E samples [nrec, N, Bpad] (Bpad = B rounded up to
 * NREM_TILE_SIMS, simulation fastest) and the final state [3, N, Bpad].                        */
int nrem_sweep_integrate_f32(const nrem_wc_params* p, int kernel, const double* CM, const double* mapG,
                             const double* mapS, const double* G0, const double* dG, const double* sigma0,
                             const double* dsigma, const int32_t* h_map_id, const uint64_t* streams, int B,
                             int n_maps, int64_t nrec, float* E_samples, float* final_state, void* stream);

/* Same with a per-node parameter table (node_params: NULL or device [NREM_NODE_PARAMS, N], node-lane kernels only). */
int nrem_sweep_integrate_f32_ex(const nrem_wc_params* p, int kernel, const double* CM, const double* mapG,
                                const double* mapS, const double* G0, const double* dG, const double* sigma0,
                                const double* dsigma, const int32_t* h_map_id, const uint64_t* streams,
                                const double* node_params, int B, int n_maps, int64_t nrec, float* E_samples,
                                float* final_state, void* stream);

/* Large connectomes (BASELINE configs[4]; netwWilsonCowanPlastic.py:64-68 allows any nnodes): integrates B simulations of
 * 16 <= nnodes <= 8192 nodes with one launch per Euler step (tcgen05 GEMM of the whole batch with the node update fused
 * onto the TMEM accumulator, wc_big.cuh).  Same noise stream, a_ie treatment and outputs as nrem_sweep_integrate_f32;
 * mapG/mapS are ONE optional per-node map each (device [N], NULL = ones).  All arrays are device pointers.
 * E_samples [nrec, N, Bpad] (may be NULL), final_state [3, N, Bpad], coup_first (may be NULL) receives the coupling
 * SC.E of the first step [N, Bpad].
 * kernel: precision of the contraction: 2 = one TF32 pass, 3 = 3xTF32, 4 = TF32 + two BF16 correction passes ("tcb"),
 *         7 = 3xBF16 ("bf3": operands split into two bf16 each, ~4e-6 per product), 0 = auto (= 7).                  */
int nrem_big_integrate_f32(const nrem_wc_params* p, int kernel, const double* CM, const double* mapG, const double* mapS,
                           const double* G0, const double* dG, const double* sigma0, const double* dsigma,
                           const uint64_t* streams, int B, int64_t nrec, float* E_samples, float* final_state,
                           float* coup_first, void* stream);
/* Same with every node parameter as a per-node vector (netwWilsonCowanPlastic.py:21): node_params = NULL or device
 * [NREM_NODE_PARAMS, N] in the order of nrem_wc_run_f64_ex; kernel 0 / 7 (bf3) only.                                  */
int nrem_big_integrate_f32_ex(const nrem_wc_params* p, int kernel, const double* CM, const double* mapG, const double* mapS,
                              const double* G0, const double* dG, const double* sigma0, const double* dsigma,
                              const uint64_t* streams, const double* node_params, int B, int64_t nrec, float* E_samples,
                              float* final_state, float* coup_first, void* stream);

/* Self-test of the tcgen05 contraction used by kernels 2/3: out[128,96] = E[128,96] x SCp[96,96]^T (float32,
 * device pointers).  passes = 1 (TF32) or 3 (3xTF32).  The shared-memory descriptor fields (bytes) and the
 * instruction descriptor can be overridden for diagnosis; 0 selects the library's own values.             */
int nrem_selftest_tc_coupling(const float* E, const float* SCp, float* out, int passes, uint32_t lboA, uint32_t sboA,
                              uint32_t lboB, uint32_t sboB, uint32_t idesc, void* stream);

/* Measures the FP32 FMA-pipe peak of the current device (register-only FMA chains on every SM, best of 3,
 * ~20 ms each): the roofline denominator of the integrator.  Blocking.  h_tflops, h_ms: host.       */
int nrem_measure_fma_peak(double* h_tflops, double* h_ms);

/* Device time (CUDA events) of the integrator launches of the last nrem_sweep_integrate_f32 call on this thread. */
double nrem_last_integrate_ms(void);

/* Kernel launches issued by this library on the calling thread since the last reset. */
int64_t nrem_launch_count(int reset);

#ifdef __cplusplus
}
#endif
#endif /* NREMFC_H */

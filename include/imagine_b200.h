/*
 * imagine_b200 -- C ABI of the B200-native IMAGINE hot path.
 *
 * One shared library, `libimagine_b200.so`, plain C linkage, plain pointers and sizes.
 * The reference (IMAGINE-Consortium/imagine v2.0.0a3) is pure Python and has no FFI for
 * this path; each entry point below cites the reference interface it replaces
 * (file:line relative to the reference root).  INTEGRATION.md shows the ctypes stub a
 * maintainer adds on the reference side.
 *
 * Conventions
 *  - every `double*` / `int32_t*` buffer argument is DEVICE memory owned by the caller,
 *    unless the parameter is documented as HOST;
 *  - `stream` is a cudaStream_t passed as void* (NULL = default stream); all work is
 *    enqueued on it; functions documented as "synchronises" wait for their result;
 *  - every function returns 0 (IMB_OK) or an IMB_E* code; `imb_strerror` explains it;
 *    nothing exits or throws across the boundary;
 *  - all arithmetic is IEEE double; canonical units kpc, microgauss, cm^-3, rad, GHz, K;
 *  - the library owns only an internal scratch workspace (grown on demand, released by
 *    `imb_finalize`); one host thread per device.
 */
#ifndef IMAGINE_B200_H
#define IMAGINE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IMB_VERSION 100

/* ---- error codes ------------------------------------------------------------------ */
enum {
  IMB_OK = 0,
  IMB_EINVAL = 1,        /* bad argument (AssertionError / ValueError on the Python side) */
  IMB_ECUDA = 2,         /* CUDA runtime error; imb_last_cuda_error() has the text        */
  IMB_ENOMEM = 3,        /* workspace allocation failed                                   */
  IMB_ENOTINIT = 4,      /* imb_init not called                                           */
  IMB_EBOUNDS = 5,       /* tabulated point outside the grid hull (ValueError)            */
  IMB_ESINGULAR = 6,     /* exactly singular matrix in LU (numpy.linalg.LinAlgError)      */
  IMB_EUNSUPPORTED = 7,  /* combination not implemented                                   */
  IMB_ESTATE = 8         /* imb_gram_phase2/3 after another call used the library's scratch workspace:
                            the phases of one sharded evaluation must follow each other on one stream   */
};

/* ---- geometry: UniformGrid, cartesian (fields/grid.py:189-275) --------------------- */
typedef struct {
  double lo[3];     /* box lower limits x,y,z [kpc]                                       */
  double hi[3];     /* box upper limits                                                   */
  int32_t n[3];     /* vertex counts; vertex i at lo + i*(hi-lo)/(n-1), n==1 -> lo        */
  int32_t reserved;
} imb_grid;

/* ---- field models (fields/basic_fields.py, fields/test_field.py) ------------------- */
enum {
  IMB_FIELD_CONST_B = 0,      /* p = Bx,By,Bz [uG]              basic_fields.py:29-40     */
  IMB_FIELD_GAUSS_B = 1,      /* p = a0,b0 [uG], stochastic     test_field.py:84-96       */
  IMB_FIELD_CONST_NE = 2,     /* p = ne [cm^-3]                 basic_fields.py:55-56     */
  IMB_FIELD_EXP_NE = 3,       /* p = n0,Re,he                   basic_fields.py:81-88     */
  IMB_FIELD_COS_NE = 4,       /* p = n0,a,alpha,b,beta,c,gamma  test_field.py:36-43       */
  IMB_FIELD_RANDOM_NE = 5,    /* p = mean,std,min_ne, stochastic basic_fields.py:111-127  */
  IMB_FIELD_CRE_ANALYTIC = 6, /* p = j0,r0,z0,r_sun   (new: base_fields.py:104-113 stub;  */
  IMB_FIELD_CRE_CONST = 7,    /* p = j0                parameter names hamx/cre_analytic) */
  IMB_FIELD_LSA_B = 8,        /* p = b0 [uG], psi0, psi1, chi0 [rad], r0, z0, r_min, r_max [kpc]:  */
                              /* WMAP-3yr logarithmic spiral (parameter names hamx/breg_lsa.py:20-31) */
  IMB_FIELD_JF12_B = 9        /* p = b_arm1..7, b_ring [uG], h_disk, w_disk [kpc], Bn, Bs [uG], rn, rs, */
                              /* wh, z0 [kpc], B0_X [uG], Xtheta [deg], rpc_X, r0_X [kpc]: Jansson &    */
                              /* Farrar 2012 regular field (names and order of hamx/breg_jf12.py:17-39) */
};
#define IMB_MAX_FIELDS 16

typedef struct {
  int32_t kind;     /* IMB_FIELD_*                                                        */
  int32_t stream;   /* Philox stream id: decorrelates stochastic fields sharing a seed    */
  int32_t seed_row; /* row of the `seeds` table this field draws from (each FieldFactory  */
  int32_t reserved; /* owns its own seed list, pipelines/pipeline.py:805-813)             */
  double p[24];     /* parameters in canonical units (JF12 takes 20)                        */
} imb_field_desc;

/* ---- line-of-sight integration ----------------------------------------------------- */
typedef struct {
  double observer[3];  /* Galactic-centric observer position [kpc] (test_ham.xml:26-30)   */
  double dr;           /* radial step oc_r_res [kpc]                                      */
  int32_t n_steps;     /* samples at r_j = (j+1/2) dr, j = 0..n_steps-1                   */
  int32_t out_cols_per_dest; /* 0: outputs are row-major [n_real, m].  > 0 (must divide m): pixel-  */
                       /* sharded layout, pixel p of realisation r at                      */
                       /* (p / cols) * out_dest_stride + r * cols + p % cols               */
  double nu_ghz;       /* synchrotron frequency [GHz]                                     */
  double alpha, beta, theta; /* CRE spectral index p = alpha - beta*R - theta*z           */
  double e0_gev;       /* CRE reference energy [GeV]                                      */
  double sync_k;       /* K(alpha,nu) computed by the host when beta == theta == 0        */
  int64_t out_dest_stride;   /* doubles between destination blocks (see out_cols_per_dest)  */
  /* Fused exchange (optional): with n_peers = m / out_cols_per_dest > 0 the block of destination */
  /* d is stored straight into rank d's buffer over NVLink peer memory: the out_* pointers address */
  /* MY slot inside my own (symmetric) receive buffer and peer_delta[d] = (rank d's buffer base -  */
  /* mine) in doubles is added for pixels owned by d.  out_dest_stride is ignored.                 */
  int32_t n_peers;
  int32_t reserved;
  int64_t peer_delta[16];
} imb_los_cfg;

/* ---- likelihood options ------------------------------------------------------------ */
enum { IMB_COV_NONE = 0, IMB_COV_VARIANCE = 1, IMB_COV_DENSE = 2 };
enum { IMB_EST_EMPIRICAL = 0, IMB_EST_OAS = 1, IMB_EST_DIAGONAL = 2 };
enum { IMB_ROUTE_AUTO = 0, IMB_ROUTE_GRAM = 1, IMB_ROUTE_DENSE = 2 };
/* out_info[0] bit field */
enum {
  IMB_INFO_ROUTE_GRAM = 1, IMB_INFO_ROUTE_DENSE = 2, IMB_INFO_CHOLESKY = 4,
  IMB_INFO_LU = 8, IMB_INFO_NEGATIVE_DET = 16
};

/* ---- life cycle -------------------------------------------------------------------- */
int imb_init(int device);
int imb_finalize(void);
int imb_version(void);
const char* imb_strerror(int code);
const char* imb_last_cuda_error(void);
/* number of kernels this library has launched since imb_init (bench: gpu_launches) */
int64_t imb_launch_count(void);

/* ---- stage 1: fields --------------------------------------------------------------- */
/*
 * Evaluates every field of `fields` on `grid` and sums same-type fields left to right.
 * Replaces, for all realisations at once: Field.get_data (fields/field.py:135-173),
 * the compute_field bodies cited above and the same-type sum in
 * Simulator.prepare_fields (simulators/simulator.py:166-178).
 *   seeds      HOST [n_seed_rows, n_real] ensemble seeds (Field.ensemble_seeds), one row per
 *              seed list; fields[f].seed_row selects the row of a stochastic field
 *   out_B      [rows_B, ncell, b_stride]  b_stride = 3 (reference layout, component
 *              innermost) or 4 (padded, 4th component = 0); rows_* is n_real if a
 *              stochastic field of that type is present, else 1 (IMB_EINVAL otherwise).
 *              Realisation-interleaved layouts (need rows_B == n_real, 32-byte aligned):
 *              b_stride = 96 | 48 | 24: [ceil(n_real/RL), ncell, 3, RL] with RL = 32 | 16 | 8 --
 *              realisation r lives in group r/RL, lane r%RL, and the RL realisations of one
 *              component of one vertex are contiguous: what the lane-per-realisation LOS kernel
 *              reads with one coalesced 64-bit load per corner component (fastest; pick the
 *              widest RL that divides n_real, else 24 with a padded last group).  Lanes beyond
 *              n_real are written as zeros.
 *   out_ne     [rows_ne, ncell]     out_cre  [rows_cre, ncell]
 * A NULL output skips that field type.  Asynchronous.
 */
int imb_field_eval(const imb_grid* grid, const imb_field_desc* fields, int nfields,
                   const uint64_t* seeds, int n_seed_rows, int n_real,
                   double* out_B, int rows_B, int b_stride,
                   double* out_ne, int rows_ne,
                   double* out_cre, int rows_cre, void* stream);

/*
 * The magnetic field of ONE parameter point written into a slice of the lanes of a wider interleaved
 * buffer, so that several parameter points share one buffer and ONE imb_los_integrate call marches
 * them together (the rays, cells and trilinear weights do not depend on the parameter point; with n_local
 * = 8 realisations per GPU and point, four points fill the 32 lanes of the fastest ray-march variant).
 *   b_stride       24 | 48 | 96: lanes per group of THIS call (RL = b_stride / 3 >= n_real, one group)
 *   out_b_stride   48 | 96 >= b_stride: the buffer is [1, ncell, 3, out_b_stride / 3]
 *   lane_offset    first lane of this point (multiple of 8, lane_offset + RL <= out_b_stride / 3)
 * Same arithmetic, same Philox keys as imb_field_eval: lane lane_offset + r holds realisation r bit for
 * bit.  Only magnetic-type entries of `fields` are evaluated.  Asynchronous.
 * (batched counterpart of Pipeline._core_likelihood over several cubes, pipelines/pipeline.py:976-979)
 */
int imb_field_eval_lanes(const imb_grid* grid, const imb_field_desc* fields, int nfields,
                         const uint64_t* seeds, int n_seed_rows, int n_real,
                         double* out_B, int b_stride, int out_b_stride, int lane_offset, void* stream);

/* rows a field type needs: n_real if any stochastic field of `type` (0 B, 1 ne, 2 cre)
 * is in `fields`, 1 if only deterministic ones, 0 if none. HOST only, no CUDA work. */
int imb_field_rows(const imb_field_desc* fields, int nfields, int type, int n_real);

/*
 * TestSimulator.simulate (simulators/test_simulator.py:47-69): out[r,p] =
 * scale * B[r, cell(p), b_axis] * ne[r, cell(p)], nearest vertex per axis (ties to the
 * lower vertex), for all realisations.  points [P,3] kpc.  Synchronises; returns
 * IMB_EBOUNDS if any point is outside the grid hull (scipy raises ValueError).
 */
int imb_test_simulate(const imb_grid* grid, const double* B, int rows_B, int b_stride,
                      const double* ne, int rows_ne, int n_real, int b_axis,
                      const double* points, int n_points, double scale,
                      double* out, void* stream);

/* ---- stage 2: line of sight -------------------------------------------------------- */
/* RING-ordered HEALPix pixel-centre unit vectors, dirs [12*nside^2, 3].
 * (map contract: observables/dataset.py:225-232, simulators/hammurabi.py:139-153) */
int imb_healpix_ring_dirs(int nside, double* dirs, void* stream);

/*
 * Integrates all realisations along the directions `dirs` [m,3] (unit vectors in the
 * Galactic-centric frame).  Replaces Hammurabi.simulate + the hammurabiX run behind it
 * (simulators/hammurabi.py:193-227); numeric definition: oracle/los.py.
 *   B   [rows_B, ncell, b_stride] (b_stride 3 | 4), or one of the interleaved layouts of
 *       imb_field_eval (b_stride 24 | 48 | 96);   ne [rows_ne, ncell] or NULL   cre [rows_cre, ncell] or NULL
 *   out_dm, out_fd, out_I, out_Q, out_U  each [n_real, m] or NULL
 * fd/dm need ne; I/Q/U need cre (IMB_EINVAL otherwise).  Asynchronous.
 */
int imb_los_integrate(const imb_grid* grid, const imb_los_cfg* cfg,
                      const double* dirs, int m,
                      const double* B, int rows_B, int b_stride,
                      const double* ne, int rows_ne,
                      const double* cre, int rows_cre, int n_real,
                      double* out_dm, double* out_fd,
                      double* out_I, double* out_Q, double* out_U, void* stream);



/* PI = sqrt(Q^2+U^2), PA = atan2(U,Q)/2 (keys 'PI','PA', observable_dict.py:15-20) */
int imb_pol_derive(const double* Q, const double* U, int64_t count,
                   double* out_PI, double* out_PA, void* stream);

/* ---- stage 3: ensemble statistics and log-likelihood -------------------------------- */
/* pmean (tools/parallel_ops.py:45-53): mean[i] = (sum_r sims[r,i]) / N, r ascending.   */
int imb_ensemble_mean(const double* sims, int N, int m, double* mean, void* stream);
/* pvar (parallel_ops.py:57-66): population variance over realisations.                  */
int imb_ensemble_var(const double* sims, int N, int m, double* var, void* stream);

/*
 * empirical_mcov / oas_mcov / diagonal_mcov (tools/covariance_estimator.py:48-237):
 * mean [m] and dense covariance [m,m]; out_rho (HOST, may be NULL) receives the OAS
 * shrinkage coefficient.  Synchronises.
 */
int imb_mcov(const double* sims, int N, int m, int estimator,
             double* out_mean, double* out_cov, double* out_rho, void* stream);

/*
 * One observable key of EnsembleLikelihood.call (likelihoods/ensemble_likelihood.py:68-107):
 *   -1/2 [ d^T A^-1 d + sign * log|2 pi A| ],  A = meas_cov + cov_est(sims),  d = meas - mean
 *   sims [N,m]; meas [m]; cov: [m] variance (cov_kind 1) or [m,m] dense (cov_kind 2)
 *   route: GRAM needs cov_kind == VARIANCE and estimator in {EMPIRICAL, OAS}; AUTO picks
 *   GRAM when legal.  DENSE forms A and factorises it: Cholesky when A is symmetric
 *   positive definite, partially pivoted LU otherwise (the reference is LU throughout and
 *   its own tests feed non-SPD matrices, tests/test_likelihood.py:63-113).
 *   out_loglike HOST [1]; out_info HOST [2] (may be NULL): [0] IMB_INFO_* bits,
 *   [1] LAPACK-style info.  Synchronises.
 */
int imb_loglike_ensemble(const double* sims, int N, int m, const double* meas,
                         const double* cov, int cov_kind, int estimator, int route,
                         int use_trace_approximation,
                         double* out_loglike, int32_t* out_info, void* stream);

/*
 * The loop over observables of EnsembleLikelihood.call (ensemble_likelihood.py:86-105) in one
 * batched pass: K keys with the same [N, m], diagonal measurement covariance, estimator
 * EMPIRICAL or OAS, Gram route (a key whose shrunk diagonal is not positive falls back to the
 * dense LU route on its own).  sims/meas/var are HOST arrays of K DEVICE pointers;
 * out_loglike HOST [K]; out_info HOST [2K] or NULL.  One synchronisation for all keys.
 */
int imb_loglike_ensemble_batch(int K, const double* const* sims, int N, int m,
                               const double* const* meas, const double* const* var,
                               int estimator, double* out_loglike, int32_t* out_info,
                               void* stream);

/*
 * The same Gram route in three phases for an ensemble whose PIXELS are sharded over ranks (each
 * rank holds all N realisations of m_local of the m_total pixels of every key, e.g. after an
 * all-to-all of the per-rank map blocks).  Row r of key k lives at
 *   sims[k] + (r / rows_per_block) * block_stride + (r % rows_per_block) * row_stride.
 *   phase 1: G[K,N,N] (device, caller-owned) <- partial Gram of the local pixels.  SUM-all-reduce G.
 *   phase 2: P2[K, N*N+N+4] <- partial weighted Gram, y, diagonal sums (needs the reduced G and
 *            m_total).  SUM-all-reduce P2.
 *   phase 3: K = I + c Gw, Cholesky, log-likelihood per key -> out_loglike HOST [K]; every rank
 *            gets the identical number.  out_info[2k+1] = -1 flags a non-positive shrunk diagonal
 *            (fall back to imb_loglike_ensemble on gathered maps).  Synchronises.
 * The three calls of one evaluation must not be interleaved with other library calls (they share
 * the scratch workspace).  K <= 16.  meas/var point at the LOCAL pixel slices.
 */
int imb_gram_phase1(int K, const double* const* sims, int N, int m_local, int rows_per_block,
                    int64_t block_stride, int64_t row_stride, const double* const* meas, double* G,
                    void* stream);
int imb_gram_phase2(int K, const double* const* sims, int N, int m_local, int rows_per_block,
                    int64_t block_stride, int64_t row_stride, const double* const* meas,
                    const double* const* var, int64_t m_total, int estimator, double* G, double* P2,
                    void* stream);
int imb_gram_phase3(int K, int N, int m_local, double* G, double* P2, double* out_loglike,
                    int32_t* out_info, void* stream);

/*
 * ---- collectives behind the C ABI (csrc/comm.cu) --------------------------------------------------
 * For hosts without torch.distributed (the reference's MPI mode: tools/mpi_helper.py, pipelines/
 * pipeline.py:945-996).  NCCL is dlopen'ed on first use.  Rank 0 calls imb_comm_unique_id, the host
 * ships the 128 bytes to every rank (MPI_Bcast, file, socket), every rank calls imb_comm_init after
 * imb_init(device).  imb_allreduce_sum is the in-place SUM all-reduce of doubles on `stream`
 * (mpi_helper.py:132).  imb_loglike_ensemble_sharded = imb_gram_phase1 -> all-reduce(G) ->
 * imb_gram_phase2 -> all-reduce(P2) -> imb_gram_phase3 in ONE call: EnsembleLikelihood.call
 * (likelihoods/ensemble_likelihood.py:68-107) for K keys whose pixels are sharded over the ranks;
 * arguments as the phase calls; G [K,N,N] and P2 [K,N*N+N+4] are caller-owned device scratch; the
 * values are replicated on every rank.  IMB_ENOTINIT without a communicator, IMB_EUNSUPPORTED when
 * libnccl.so.2 cannot be loaded.
 */
int imb_comm_unique_id(void* out128);
int imb_comm_init(int nranks, int rank, const void* unique_id128);
int imb_comm_finalize(void);
int imb_comm_size(void);
int imb_comm_rank(void);
int imb_allreduce_sum(double* device_buf, int64_t count, void* stream);
/* All-to-all of doubles on the library's communicator: block r (count_per_rank doubles) of `send` goes to
 * rank r, block q of `recv` arrives from rank q; `recv` must not alias `send`.  The exchange of the
 * destination-major map buffer the ray march writes (imb_los_cfg.out_cols_per_dest): afterwards every rank
 * holds all realisations of its pixels (mpi_helper.py's distributed arrays, tools/mpi_helper.py:25-90).
 * Asynchronous on `stream`. */
int imb_alltoall(const double* send, double* recv, int64_t count_per_rank, void* stream);

int imb_loglike_ensemble_sharded(int K, const double* const* sims, int N, int m_local,
                                 int rows_per_block, int64_t block_stride, int64_t row_stride,
                                 const double* const* meas, const double* const* var, int64_t m_total,
                                 int estimator, double* G, double* P2, double* out_loglike,
                                 int32_t* out_info, void* stream);

/*
 * Asynchronous variants (K <= 16): the per-key values land in DEVICE memory (out_device [K]) with
 * flag_device [K] (may be NULL; 0 = ok, non-zero = Gram route not applicable for that key) and the
 * call does not synchronise.  Lets a caller pipeline many parameter points -- UltraNest
 * `vectorized`, emcee walkers, the MPI loop over cubes (pipelines/pipeline.py:976-979) -- and read
 * all values back once.  No dense-route fallback here: check the flags.
 */
int imb_loglike_ensemble_batch_async(int K, const double* const* sims, int N, int m,
                                     const double* const* meas, const double* const* var,
                                     int estimator, double* out_device, int32_t* flag_device,
                                     void* stream);
int imb_gram_phase3_async(int K, int N, int m_local, double* G, double* P2, double* out_device,
                          int32_t* flag_device, void* stream);

/* EnsembleLikelihoodDiagonal.call (likelihoods/ensemble_likelihood.py:137-173);
 * var [m] = diagonal of the measurement covariance.  Synchronises. */
int imb_loglike_ensemble_diagonal(const double* sims, int N, int m, const double* meas,
                                  const double* var, double* out_loglike, void* stream);

/* SimpleLikelihood.call, one key (likelihoods/simple_likelihood.py:33-84).  Synchronises. */
int imb_loglike_simple(const double* sims, int N, int m, const double* meas,
                       const double* cov, int cov_kind,
                       double* out_loglike, int32_t* out_info, void* stream);

/* mask_obs / mask_var (tools/masker.py:25-69): out[r, k] = in[r, idx[k]], idx [m_out].  */
int imb_gather_columns(const double* in, int rows, int m, const int32_t* idx, int m_out,
                       double* out, void* stream);
/* mask_cov (tools/masker.py:72-116): out[a,b] = cov[idx[a], idx[b]].                     */
int imb_gather_cov(const double* cov, int m, const int32_t* idx, int m_out,
                   double* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* IMAGINE_B200_H */

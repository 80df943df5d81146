/* qcd_b200.h — C-ABI of libqcd_b200.so: the batched CircuitDesigner hot path on B200 (sm_100a).
 *
 * The reference (philippaltmann/QCD, qcd-gym 0.2.0) is pure Python and has no FFI of its own.
 * Each entry point below replaces a Python-level interface of /root/reference/qcd_gym/env.py;
 * the binding a maintainer would add (a ctypes stub inside qcd_gym/env.py) is in INTEGRATION.md.
 *
 * Conventions
 *   - every pointer inside qcd_batch_t is a DEVICE pointer (tensor.data_ptr()); the struct itself
 *     lives in host memory and is read during the call only (it may be reused / mutated afterwards);
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream);
 *   - calls are asynchronous on `stream`, never synchronise (except qcd_step_sync), never allocate, and
 *     keep no per-call state: the only process-wide state is idempotent and thread-safe (the QCD_PDL
 *     environment switch read once; one "shared-memory opt-in done" bit per kernel and device);
 *   - any of the per-step OUTPUT pointers, `obs` and the action pointer may also be pinned host memory
 *     mapped into the device address space (cudaHostAlloc / UVA): the kernels then read the action and
 *     write the results over PCIe directly (the single-env path does this, see qcd_step_sync);
 *   - return value: 0 ok, <0 invalid argument (QCD_E_*), >0 a cudaError_t from the launch;
 *     a batch with n_envs == 0 is a valid no-op;
 *   - the `_host` variants take HOST pointers for actions/results and issue the copies on `stream`.
 *
 * Layouts (N = n_envs, eta = n_qubits, D = 2^eta, S = D for SP, D*D for UC):
 *   state      complex128 [N,S]   interleaved (re,im); UC row-major V[r*D+c]; qubit k = bit k of the
 *                                 basis index (little-endian, qiskit), UC gates act on the ROW index
 *   target     complex128 [S] shared, or [N,S] with QCD_F_TARGET_PER_ENV
 *   meta       uint8 [N,16]       bytes 0..11 per-qubit depth counters (QuantumCircuit.depth stack),
 *                                 bytes 12..13 uint16 operations, bytes 14..15 uint16 used-wire mask
 *   last       float64 [N,2]      last_reward,last_cost of env.py:102-104 (NULL allowed when sparse)
 *   ep_return  float64 [N], ep_length int32 [N]: running Monitor accumulators (monitor.py:37-38), NULL ok
 *   actions    [N,4] float64 or float32 (action_dtype), <o,q,c,Phi> of env.py:90-91
 *   obs        float32, row e starts at obs + e*obs_stride, holds [Re state (S) | Im state (S)]
 *              (env.py:10 `flat`); with obs_stride = 4*S the caller pre-fills the constant target
 *              half once and each row is exactly the reference observation (env.py:85)
 *   final_obs  same layout/stride as obs; written only for envs that finished this step when
 *              autoreset is on (the terminal observation); qcd_replay writes it at every reset of the
 *              tape, so the buffer ends up as T qcd_step calls leave it; NULL = do not write
 */
#ifndef QCD_B200_H
#define QCD_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QCD_ABI_VERSION 6

#define QCD_MAX_QUBITS_SP 12   /* S = 4096 complex128 = 64 KiB per env */
#define QCD_MAX_QUBITS_UC 6    /* 64x64 unitary      = 64 KiB per env */
#define QCD_MAX_DEPTH     255  /* per-qubit depth counters are uint8  */

/* qcd_batch_t.flags */
#define QCD_F_PUNISH          1u   /* env.py:135  reward -= cost                               */
#define QCD_F_SPARSE          2u   /* env.py:134  zero reward/cost unless terminated|truncated */
#define QCD_F_AUTORESET       4u   /* VectorEnv: reset finished envs inside the same launch    */
#define QCD_F_TARGET_PER_ENV  8u   /* target is [N,S] instead of shared [S]                    */
#define QCD_F_ACTIONS_STABLE 16u   /* the caller guarantees that the action buffer passed to qcd_step was
                                      completely written BEFORE the previous kernel of the stream was
                                      launched (pre-generated tapes).  qcd_step is launched with programmatic
                                      stream serialization; with this flag it reads and decodes the action
                                      before griddepcontrol.wait (overlapping the previous step's tail),
                                      without it only after the wait, which is safe when the previous kernel
                                      of the stream (a policy / sampler) produced the actions.            */

/* objective */
#define QCD_OBJ_SP 0               /* state preparation: fidelity |<psi|target>|^2  (env.py:109) */
#define QCD_OBJ_UC 1               /* unitary composition: 1-2*atan(||U-V||_F)/pi   (env.py:111) */

/* action_dtype */
#define QCD_ACT_F64 0
#define QCD_ACT_F32 1

/* per-env status written by the step / replay kernels (reference: AssertionError) */
#define QCD_ST_OK        0
#define QCD_ST_BAD_WIRE  1         /* env.py:94  wire/control outside range(eta) */
#define QCD_ST_BAD_GATE  2         /* env.py:100 gate index not in {0,1,2}        */

/* error codes (<0) */
#define QCD_E_NULL     (-1)        /* required pointer is NULL            */
#define QCD_E_RANGE    (-2)        /* n_qubits / max_depth / n_envs range */
#define QCD_E_OBJ      (-3)        /* unknown objective / dtype           */
#define QCD_E_STRIDE   (-4)        /* obs_stride < 2*S                    */

typedef struct qcd_batch {
  /* ---- configuration: CircuitDesigner.__init__ kwargs (env.py:29-40) ---- */
  int32_t  n_envs;
  int32_t  n_qubits;        /* max_qubits (eta)                                           */
  int32_t  objective;       /* QCD_OBJ_*                                                  */
  int32_t  max_depth;       /* delta; truncation depth >= delta (env.py:129)              */
  uint32_t flags;           /* QCD_F_*                                                    */
  int32_t  reserved0;
  double   depth_free;      /* delta/3       as computed by the host in float64 (env.py:112) */
  double   cost_denom;      /* delta/2*3     as computed by the host in float64 (env.py:112) */
  int64_t  obs_stride;      /* floats between consecutive obs rows, >= 2*S                */
  /* ---- persistent per-env state (read + written) ---- */
  void*    state;           /* complex128 [N,S]                                           */
  const void* target;       /* complex128 [S] or [N,S]                                    */
  void*    meta;            /* uint8 [N,16]                                               */
  double*  last;            /* float64 [N,2] or NULL (required when not sparse)           */
  double*  ep_return;       /* float64 [N] or NULL                                        */
  int32_t* ep_length;       /* int32 [N] or NULL (must be NULL iff ep_return is NULL)     */
  /* ---- per-step outputs (written) ---- */
  float*   obs;             /* float32 rows, may be NULL (observations disabled)          */
  float*   final_obs;       /* float32 rows or NULL                                       */
  double*  reward;          /* [N] env.py:136 return value                                */
  double*  metric;          /* [N] info['metric'] (env.py:133)                            */
  double*  cost;            /* [N] info['cost']                                           */
  int32_t* depth;           /* [N] info['depth']                                          */
  int32_t* operations;      /* [N] info['operations']                                     */
  int32_t* used_wires;      /* [N] info['used_wires']                                     */
  uint8_t* terminated;      /* [N] */
  uint8_t* truncated;       /* [N] */
  int8_t*  status;          /* [N] QCD_ST_*                                               */
  double*  episode_return;  /* [N] or NULL: Monitor 'r' latched when the episode finishes */
  int32_t* episode_length;  /* [N] or NULL: Monitor 'l' latched when the episode finishes */
  double*  stats;           /* [QCD_STATS_COPIES][QCD_STATS_STRIDE] or NULL: running episode statistics,
                               accumulated (+=) by qcd_step / qcd_replay themselves (fused; no extra
                               launch).  Field k of the statistics = sum over copies of stats[c][k]    */
} qcd_batch_t;

/* Per-step outputs of a replay over T steps; every pointer may be NULL (not recorded).
 * Layout [T,N] (step-major) so that row t equals what qcd_step would have written at step t. */
typedef struct qcd_tape_out {
  double*  reward;
  double*  metric;
  double*  cost;
  int32_t* depth;
  uint8_t* terminated;
  uint8_t* truncated;
  int8_t*  status;
  int32_t* operations;      /* info['operations'] of every step */
  int32_t* used_wires;      /* info['used_wires'] of every step */
} qcd_tape_out_t;

/* Episode statistics accumulated by qcd_reduce_stats: the Monitor fields of monitor.py:37-48 and the
 * termination histogram of algorithm/algorithm.py:112-114, summed over finished episodes. */
#define QCD_STATS_LEN 12
#define QCD_STATS_COPIES 64      /* replicas of the vector, one 128-byte line each (atomic spreading) */
#define QCD_STATS_STRIDE 16      /* doubles between replicas                                          */
enum {
  QCD_STAT_EPISODES = 0, QCD_STAT_SUM_R, QCD_STAT_SUM_L, QCD_STAT_SUM_D, QCD_STAT_SUM_O,
  QCD_STAT_SUM_Q, QCD_STAT_SUM_M, QCD_STAT_SUM_C, QCD_STAT_N_DONE, QCD_STAT_N_DEPTH,
  QCD_STAT_STEPS, QCD_STAT_BAD_ACTIONS
};

int qcd_abi_version(void);

/* Replaces CircuitDesigner.reset (env.py:117-121) for every env of the batch (mask == NULL) or for
 * the envs with mask[e] != 0: state <- e0 / I, counters and accumulators zeroed, obs rows written. */
int qcd_reset(const qcd_batch_t* b, const uint8_t* mask, void* stream);

/* Replaces CircuitDesigner.step (env.py:124-136 incl. _operation :90-100, _state :80-87,
 * _reward :106-114) for all N envs in ONE launch: decode, gate, depth, metric, cost, reward,
 * float32 observation, optional same-step auto-reset. */
int qcd_step(const qcd_batch_t* b, const void* actions, int action_dtype, void* stream);

/* T consecutive qcd_step calls fused into one launch with each env's state resident in shared
 * memory (the reference re-simulates whole circuits per evaluation: baselines/evo/run.py:17-33,
 * env.py:83-84).  actions is [T,N,4].  Bit-identical to T calls of qcd_step on the same batch;
 * `b`'s per-step outputs hold the values of the LAST step, `out` (optional) records every step. */
int qcd_replay(const qcd_batch_t* b, const void* actions, int action_dtype, int32_t n_steps,
               const qcd_tape_out_t* out, void* stream);

/* Adds this step's finished-episode statistics (see QCD_STAT_*) to stats[QCD_STATS_COPIES]
 * [QCD_STATS_STRIDE] (float64, device; sum over the copies = the statistics).  Reads only the per-step outputs of `b`.  Stand-alone variant of the accumulation that
 * qcd_step / qcd_replay fuse when b->stats != NULL (do not use both on the same vector). */
int qcd_reduce_stats(const qcd_batch_t* b, double* stats, void* stream);

/* Folds the QCD_STATS_COPIES replicas into out[QCD_STATS_STRIDE] (device, float64; entries >= QCD_STATS_LEN
 * are zero) with one small kernel on `stream`, and clears the replicas when zero_raw != 0 (start of the
 * next reporting interval).  `out` is what a rank hands to the statistics all-reduce (algorithm.py:99-114
 * only logs these numbers, so the exchange can run on a side stream while the envs keep stepping). */
int qcd_fold_stats(double* stats_raw, double* out, int zero_raw, void* stream);

/* Rolling window of the most recent finished episodes, on the device: SB3's ep_info_buffer = deque(maxlen=100)
 * of Monitor's info["episode"] records, which the reference's trainer averages (algorithm/algorithm.py:49,89-91,
 * 100; monitor.py:37-48).  Call after a step: the episodes that finished in that step are appended in env order;
 * ring is float64 [window][QCD_EP_FIELDS] = {r, l, d, o, q, m, c, env}, count[0] (int64, device) the number of
 * episodes pushed so far (the newest is at row (count-1) % window). */
#define QCD_EP_FIELDS 8
int qcd_push_episodes(const qcd_batch_t* b, double* ring, int64_t* count, int32_t window, void* stream);

/* qcd_step followed by cudaStreamSynchronize(stream) in one call: the single-env path (env.py:124-136 with
 * numpy in / numpy out), where the action and every output live in mapped pinned host memory. */
int qcd_step_sync(const qcd_batch_t* b, const void* actions, int action_dtype, void* stream);

/* env.step(action) of ONE env (b->n_envs == 1, env.py:124-136): `action` points to 4 doubles in ordinary host
 * memory; they travel to the step kernel as a kernel argument (no H2D copy, no PCIe read), the kernel writes
 * its results into b's (mapped pinned host) output buffers, and the call returns after
 * cudaStreamSynchronize(stream). */
int qcd_step1_sync(const qcd_batch_t* b, const double* action, void* stream);

/* HOST (pinned) destinations of qcd_step_host; every pointer may be NULL (field not copied back). */
typedef struct qcd_host_out {
  double*  reward;          /* [N] */
  double*  metric;          /* [N] */
  double*  cost;            /* [N] */
  int32_t* depth;           /* [N] */
  int32_t* operations;      /* [N] */
  int32_t* used_wires;      /* [N] */
  uint8_t* terminated;      /* [N] */
  uint8_t* truncated;       /* [N] */
  int8_t*  status;          /* [N] */
  float*   obs;             /* rows of 2*S floats (the state half), obs_stride floats apart */
  int64_t  obs_stride;
} qcd_host_out_t;

/* The reference's call with HOST buffers (env.step(action) -> obs, reward, terminated, truncated, info,
 * env.py:124-136): copy `actions_host` (pinned) to `actions_dev`, run qcd_step, copy the requested
 * results back into `out`'s host buffers; all on `stream`, no synchronisation (the caller syncs once, or
 * keeps two env groups in flight on two streams).  When both the device rows (b->obs_stride) and the host
 * rows (out->obs_stride) are compact (2*S floats) the observations move in one contiguous transfer. */
int qcd_step_host(const qcd_batch_t* b, const void* actions_host, void* actions_dev, int action_dtype,
                  const qcd_host_out_t* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* QCD_B200_H */

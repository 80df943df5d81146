// qcd_kernels.cu — batched CircuitDesigner hot path for B200 (sm_100a) + the C-ABI of include/qcd_b200.h.
//
// Kernels (DESIGN.md §4 has the roofline and the measured numbers of each):
//   step_warp_kernel   K1a  one env.step() of every env, S <= 128 amplitudes: a warp owns 32 envs,
//                           lane-per-env decode, butterflies by shuffle, register prefetch, PDL.
//   qcd_kernel<!STAGED> K1b the same step for S >= 256: pair-streaming, one warp = one env per CTA.
//   replay_warp_kernel K2a  T steps in one launch, S <= 256: the warp's tile stays in shared memory
//                           (cp.async.bulk in / out).
//   replay_reg_kernel  K2b  T steps in one launch, S >= 512: one CTA per env, the state in REGISTERS (16
//                           amplitudes per thread), gates by register arithmetic / shuffles, no barrier
//                           inside the step loop, metric steps finalized per chunk.
//   qcd_kernel<STAGED>      first replay version; kept as the S <= 256 path for per-env targets.
//   reset_kernel, stats_kernel, fold_stats_kernel, push_episodes_kernel
// One launch = one step (or one tape) of ALL envs; the step kernels are HBM-bound
// (32*S + 8*S + ~96 bytes per env-step, S = 2^eta (SP) or 4^eta (UC) complex128 amplitudes).
//
// Reference semantics reproduced (file:line in /root/reference):
//   qcd_gym/env.py:90-100   action decode <o,q,c,Phi>            -> decode_kind(), decode_action()
//   qcd_gym/env.py:83-84    gate applied to Statevector/Operator -> apply_pair() / mul_phase() / rot_rx()
//                            (incremental: the reference re-simulates the whole circuit, same ops in the
//                            same order)
//   qcd_gym/env.py:86,129   depth / operations / used wires, truncation -> book_gate()
//   qcd_gym/env.py:106-114  fidelity | arctan-Frobenius metric, depth cost -> finalize_step()
//   qcd_gym/env.py:134-135  sparse / punish shaping                -> finalize_step()
//   qcd_gym/env.py:10,85    float32 observation [Re | Im]          -> store phase of every kernel
//   qcd_gym/env.py:117-121  reset                                   -> fused auto-reset + reset_kernel
//   qcd_gym/wrappers/monitor.py:37-48 episode return/length        -> phase3(), stats_kernel
//
// Mapping.  A gate on qubit w only couples amplitude pairs (i0, i1 = i0 | 1<<w).  Every gate of the
// reference's set {P, RX, CP, CX} is a 2x2 butterfly on such a pair (CP/CX predicated on the control
// bit of i0); gates are heterogeneous per env for free (only the address / lane pattern and two
// doubles differ), and UC is the same butterfly on the row bits (w + eta) of the row-major unitary.
// The per-env scalar work (floor-decode, sincos, depth stack, cost, reward) is done by ONE lane per
// env with consecutive lanes on consecutive envs (coalesced scalar I/O).

#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>
#include <stdlib.h>

#include <atomic>

#include "qcd_b200.h"

namespace qcd {

// ---- compile-time tuning knobs (defaults = the values measured best on B200, profiles/r01_summary.md;
// override with QCD_NVCC_EXTRA="-DNAME=value" python -m qcd_b200.build)
#ifndef QCD_REG_STEP_MAX_LOG2S
#define QCD_REG_STEP_MAX_LOG2S 7        // S <= 128: warp-autonomous step kernel; above: pair-streaming
#endif
#ifndef QCD_STEP_IMPL
#define QCD_STEP_IMPL 2                 // small-S step path: 0 pair-streaming, 2 warp-autonomous
#endif
// warp-autonomous step kernel
#ifndef QCD_WARP_STEP_WARPS
#define QCD_WARP_STEP_WARPS 1           // warps per CTA (each warp is autonomous; 1 balances best over 148 SMs)
#endif
#ifndef QCD_WARP_STEP_MINB
#define QCD_WARP_STEP_MINB 16           // min CTAs per SM for __launch_bounds__ (caps registers at 128)
#endif
#ifndef QCD_WARP_STEP_TILE
#define QCD_WARP_STEP_TILE 32           // envs per warp tile (0 = 8 KiB of state per tile, clamped to 8..32)
#endif
#ifndef QCD_WARP_STEP_STAGE
#define QCD_WARP_STEP_STAGE 0           // 1: stage the tile in smem with cp.async.bulk (TMA) when it fits 16 KiB
#endif                                  //    (measured slower than the register prefetch, kept as an option)
#ifndef QCD_WARP_PF_V1
#define QCD_WARP_PF_V1 8                // rounds of state prefetched in registers at V = 1 amplitude per lane
#endif
#ifndef QCD_WARP_PF_V2
#define QCD_WARP_PF_V2 2                // ... at V = 2
#endif
#ifndef QCD_WARP_PF_V4
#define QCD_WARP_PF_V4 2                // ... at V = 3..4
#endif
#ifndef QCD_WARP_CHUNK
#define QCD_WARP_CHUNK 8                // rounds per metric-reduction chunk (power of two <= 32)
#endif
#ifndef QCD_WARP_UNROLL_CHUNKS
#define QCD_WARP_UNROLL_CHUNKS 0        // 1: fully unroll the chunk loop (measured slower: I-cache)
#endif
#ifndef QCD_WARP_BRANCHLESS
#define QCD_WARP_BRANCHLESS 1           // branch-free round body (select instead of switch) ...
#endif
#ifndef QCD_WARP_BRANCHLESS_MAX_V
#define QCD_WARP_BRANCHLESS_MAX_V 2     // ... for at most this many amplitudes per lane (S <= 64)
#endif
#ifndef QCD_WARP_LOG2L_S32
#define QCD_WARP_LOG2L_S32 5            // log2(lanes per env) of the warp kernels at S = 32 (fewer: slower)
#endif
#ifndef QCD_WARP_LOG2L_S64
#define QCD_WARP_LOG2L_S64 5            // ... at S = 64
#endif
// pair-streaming step kernel
#ifndef QCD_PAIR_SMALL_MAX_S
#define QCD_PAIR_SMALL_MAX_S 4096       // up to this S the CTA has QCD_PAIR_SMALL_THREADS threads
#endif
#ifndef QCD_PAIR_SMALL_THREADS
#define QCD_PAIR_SMALL_THREADS 32       // one warp (= one env) per CTA measured best at every S
#endif
#ifndef QCD_PAIR_NARROW_MAX_S
#define QCD_PAIR_NARROW_MAX_S 256       // pair-streaming step: for S up to this, only QCD_PAIR_NARROW_L lanes per env
#endif
#ifndef QCD_PAIR_NARROW_L
#define QCD_PAIR_NARROW_L 8             // (several envs per warp => several decode lanes busy)
#endif
#ifndef QCD_PAIR_L2_PREFETCH
#define QCD_PAIR_L2_PREFETCH 1          // pull the env's lines into L2 while lane 0 decodes ...
#endif
#ifndef QCD_PAIR_L2_PREFETCH_MAX_S
#define QCD_PAIR_L2_PREFETCH_MAX_S 1024 // ... for S up to this (hurts at 32-64 KiB per env)
#endif
// replay kernels
#ifndef QCD_REPLAY_WARP
#define QCD_REPLAY_WARP 1               // replay for S <= 256: warp-autonomous, tile resident in smem
#endif
#ifndef QCD_WARP_REPLAY_WARPS
#define QCD_WARP_REPLAY_WARPS 1
#endif
#ifndef QCD_WARP_REPLAY_MINB
#define QCD_WARP_REPLAY_MINB 16
#endif
#ifndef QCD_WARP_REPLAY_CHUNK
#define QCD_WARP_REPLAY_CHUNK 8          // rounds per metric-reduction chunk of replay_warp_kernel
#endif
#ifndef QCD_WARP_REPLAY_TILE_BYTES
#define QCD_WARP_REPLAY_TILE_BYTES 8192 // state bytes per warp tile (ghz5: 16 envs)
#endif
constexpr int K_NONE = 0, K_P = 1, K_RX = 2, K_CP = 3, K_CX = 4;
constexpr int ACT_IMMEDIATE = 2;   // internal action mode of the step kernels: the (single) action is a kernel argument

constexpr int RF_RESET = 1;    // episode finished and autoreset is on: state <- e0 / I after this step
constexpr int RF_METRIC = 2;   // the metric of this step is needed
constexpr int RF_LAST = 4;     // last step of the launch: write observation rows

struct __align__(16) GateRec {
  double cr, sr;   // P/CP: e^{i theta} = (cr, sr);  RX: cos(theta/2), sin(theta/2)
  int kind;
  int wbit;        // bit of the flat index the gate targets (w, or w + eta for UC)
  int cbit;        // control bit (CP, CX)
  int flags;       // RF_*
};

// per-env scalars carried in the registers of the env's control thread between decode and finalize
struct EnvCtl {
  uint4 meta;          // bytes 0..11 depth stack, .w = ops (lo16) | used mask (hi16)
  double last_m, last_c;
  double ep_ret;
  int ep_len;
  // results of decode for the current step
  double cost_raw;
  int depth, ops, used_cnt;
  int status;
  bool term, trunc;
};

__device__ __forceinline__ uint32_t get_byte(const uint4& m, int i) {
  const uint32_t word = (i < 4) ? m.x : ((i < 8) ? m.y : m.z);
  return (word >> ((i & 3) * 8)) & 0xffu;
}

__device__ __forceinline__ void set_byte(uint4& m, int i, uint32_t v) {
  const uint32_t sh = (i & 3) * 8;
  const uint32_t keep = ~(0xffu << sh);
  const uint32_t val = v << sh;
  if (i < 4) m.x = (m.x & keep) | val;
  else if (i < 8) m.y = (m.y & keep) | val;
  else m.z = (m.z & keep) | val;
}

__device__ __forceinline__ void load_action(const void* __restrict__ actions, int act_f32, long idx,
                                            double& a0, double& a1, double& a2, double& a3) {
  if (act_f32) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(actions) + idx);
    a0 = (double)v.x; a1 = (double)v.y; a2 = (double)v.z; a3 = (double)v.w;   // exact widening
  } else {
    const double2* p = reinterpret_cast<const double2*>(actions) + 2 * idx;
    const double2 lo = __ldg(p), hi = __ldg(p + 1);
    a0 = lo.x; a1 = lo.y; a2 = hi.x; a3 = hi.y;
  }
}

// env.py:90-100: the part of the decode that depends on the action alone (parallel over steps in the
// replay kernels).  Integer results are bit-exact by construction: floor() is taken in float64 on
// exactly-widened inputs and the range test is done before the cast.
struct GateKind { int status, kind, w, c; bool term; double cr, sr; };

__device__ __forceinline__ GateKind decode_kind(const int n_qubits, double a0, double a1, double a2, double a3) {
  const double fg = floor(a0), fw = floor(a1), fc = floor(a2);
  const double eta = (double)n_qubits;
  GateKind k;
  k.status = QCD_ST_OK; k.kind = K_NONE; k.w = 0; k.c = 0; k.term = false; k.cr = 1.0; k.sr = 0.0;
  if (!(fw >= 0.0 && fw < eta && fc >= 0.0 && fc < eta)) k.status = QCD_ST_BAD_WIRE;     // env.py:94
  else if (!(fg == 0.0 || fg == 1.0 || fg == 2.0)) k.status = QCD_ST_BAD_GATE;           // env.py:100
  if (k.status == QCD_ST_OK) {
    k.w = (int)fw; k.c = (int)fc;
    const int g = (int)fg;
    if (g == 2) k.term = true;                                 // env.py:99 (checked after 95-98,
    else if (k.w == k.c) k.kind = (g == 0) ? K_P : K_RX;       //  but g==2 matches none of those)
    else k.kind = (g == 0) ? K_CP : K_CX;
  }
  if (k.kind == K_P || k.kind == K_CP || k.kind == K_RX) {
    const double arg = (k.kind == K_RX) ? a3 / 2 : a3;         // RXGate: cos/sin(theta/2)
    sincos(arg, &k.sr, &k.cr);
  }
  return k;
}

__device__ __forceinline__ double depth_cost(const qcd_batch_t& b, const int depth) {
  return fmax(0.0, (double)depth - b.depth_free) / b.cost_denom;                       // env.py:112
}

// QuantumCircuit.depth/count_ops/idle_wires bookkeeping (env.py:74,77,86) + truncation (env.py:129)
// + cost (env.py:112): the sequential part (integer only, apart from the cost division).
template <bool WITH_COST = true>
__device__ __forceinline__ void book_gate(const qcd_batch_t& b, const int status, const int kind, const int w,
                                          const int c, const bool term, EnvCtl& e) {
  uint32_t ops = e.meta.w & 0xffffu, used = e.meta.w >> 16;
  if (kind != K_NONE) {
    const uint32_t dw = get_byte(e.meta, w), dc = get_byte(e.meta, c);
    uint32_t lvl = max(dw, dc) + 1u;
    lvl = min(lvl, 255u);                                      // saturate (only past truncation)
    set_byte(e.meta, w, lvl);
    set_byte(e.meta, c, lvl);
    ops = min(ops + 1u, 0xffffu);
    used |= (1u << w) | (1u << c);
    e.meta.w = ops | (used << 16);
  }
  uint32_t m4 = __vmaxu4(e.meta.x, __vmaxu4(e.meta.y, e.meta.z));
  m4 = __vmaxu4(m4, m4 >> 16);
  const int depth = (int)(max(m4 & 0xffu, (m4 >> 8) & 0xffu));
  e.depth = depth; e.ops = (int)ops; e.used_cnt = __popc(used);
  e.status = status; e.term = term;
  e.trunc = (status == QCD_ST_OK) &&
            (depth >= b.max_depth || (int)ops >= b.max_depth * b.n_qubits * 2);        // env.py:129
  if (WITH_COST) e.cost_raw = depth_cost(b, depth);                                    // env.py:112
}

__device__ __forceinline__ void decode_action(const qcd_batch_t& b, double a0, double a1, double a2,
                                              double a3, EnvCtl& e, GateRec& rec) {
  const GateKind k = decode_kind(b.n_qubits, a0, a1, a2, a3);
  book_gate(b, k.status, k.kind, k.w, k.c, k.term, e);
  rec.cr = k.cr; rec.sr = k.sr; rec.kind = k.kind;
  const int shift = (b.objective == QCD_OBJ_UC) ? b.n_qubits : 0;   // UC: gate acts on the row index
  rec.wbit = k.w + shift; rec.cbit = k.c + shift;
  rec.flags = ((e.term || e.trunc) && (b.flags & QCD_F_AUTORESET)) ? RF_RESET : 0;
}

// env.py:106-114 + 133-135 + Monitor accumulators.  `ra`,`rb` are the reduced metric sums.
struct StepOut { double reward, metric, cost; };

// env.py:109 / env.py:111 from the reduced sums
template <bool UC>
__device__ __forceinline__ double metric_from_sums(const double ra, const double rb) {
  if (UC) {
    const double nrm = sqrt(ra);                               // ||U - V||_F
    return 1.0 - 2.0 * atan(nrm) / CUDART_PI;                  // env.py:111
  }
  return ra * ra + rb * rb;                                    // env.py:109: abs(vdot)**2 (hypot-free:
                                                               // within 2 ulp of numpy's hypot(re,im)**2)
}

// env.py:102-104 + 133-135 + Monitor accumulators, from the raw metric of the step
__device__ __forceinline__ StepOut shape_step(const qcd_batch_t& b, EnvCtl& e, const double metric_raw) {
  StepOut o;
  if (e.status != QCD_ST_OK) {      // reference raises; batched path: env untouched, reward 0
    o.metric = metric_raw; o.cost = e.cost_raw; o.reward = 0.0;
    return o;
  }
  const bool sparse = b.flags & QCD_F_SPARSE;
  double m = metric_raw, c = e.cost_raw;
  if (!sparse) {                                               // env.py:102-104
    m = metric_raw - e.last_m; c = e.cost_raw - e.last_c;
    e.last_m = metric_raw; e.last_c = e.cost_raw;
  }
  o.metric = m; o.cost = c;                                    // env.py:133
  const bool done = e.term || e.trunc;
  double r = m, cc = c;
  if (sparse && !done) { r = 0.0; cc = 0.0; }                  // env.py:134
  if (b.flags & QCD_F_PUNISH) r -= cc;                         // env.py:135
  o.reward = r;
  e.ep_ret += r; e.ep_len += 1;                                // monitor.py:35,38
  return o;
}

template <bool UC>
__device__ __forceinline__ StepOut finalize_step(const qcd_batch_t& b, EnvCtl& e, double ra, double rb,
                                                 bool have_metric) {
  return shape_step(b, e, have_metric ? metric_from_sums<UC>(ra, rb) : CUDART_NAN);
}

__device__ __forceinline__ double2 reset_amp(bool uc, int eta, int i) {
  const bool one = uc ? ((i >> eta) == (i & ((1 << eta) - 1))) : (i == 0);
  return make_double2(one ? 1.0 : 0.0, 0.0);
}

// Gate arithmetic, written with explicit rounding intrinsics so that every kernel (pair-streaming
// step, register-resident step, shared-memory replay) produces bit-identical amplitudes.
// own * e^{i th}, e^{i th} = (cr, sr)
__device__ __forceinline__ double2 mul_phase(const double2 a, const double cr, const double sr) {
  // (.y first: it can take a.y's register once both products exist, then .x takes a.x's - no moves)
  const double p = __dmul_rn(a.y, sr), q = __dmul_rn(a.y, cr);
  const double ny = __fma_rn(a.x, sr, q);
  const double nx = __fma_rn(a.x, cr, -p);
  return make_double2(nx, ny);
}
// RX row: c*own - i*s*partner  (the same expression for both amplitudes of the pair)
__device__ __forceinline__ double2 rot_rx(const double2 own, const double2 p, const double c, const double s) {
  return make_double2(__fma_rn(c, own.x, __dmul_rn(s, p.y)), __fma_rn(c, own.y, -__dmul_rn(s, p.x)));
}

// The 2x2 butterfly on one amplitude pair; (a0,a1) at (i0, i0|1<<wbit).
__device__ __forceinline__ void apply_pair(const GateRec& g, int i0, double2& a0, double2& a1) {
  switch (g.kind) {
    case K_P:                         // diag(1, e^{i th}) on the w-bit
      a1 = mul_phase(a1, g.cr, g.sr);
      break;
    case K_CP:                        // diag(1,1,1,e^{i th}): both bits set
      if ((i0 >> g.cbit) & 1) a1 = mul_phase(a1, g.cr, g.sr);
      break;
    case K_RX: {                      // [[c, -is], [-is, c]]
      const double2 n0 = rot_rx(a0, a1, g.cr, g.sr), n1 = rot_rx(a1, a0, g.cr, g.sr);
      a0 = n0; a1 = n1;
    } break;
    case K_CX:                        // swap the pair where the control bit is set
      if ((i0 >> g.cbit) & 1) { const double2 t = a0; a0 = a1; a1 = t; }
      break;
    default: break;
  }
}

template <bool UC>
__device__ __forceinline__ void metric_acc(const double2 a, const double2 t, double& ra, double& rb) {
  if (UC) {
    const double dx = t.x - a.x, dy = t.y - a.y;
    ra += dx * dx + dy * dy;
  } else {                            // vdot: conj(psi) * target
    ra += a.x * t.x + a.y * t.y;
    rb += a.x * t.y - a.y * t.x;
  }
}

// The statistics vector is replicated QCD_STATS_COPIES times (one 128-byte line each) so that the
// per-warp atomicAdds of a launch spread over many L2 lines instead of serialising on one.
__device__ __forceinline__ double* stats_slice(const qcd_batch_t& b, const long unit) {
  return b.stats ? b.stats + (unit & (QCD_STATS_COPIES - 1)) * QCD_STATS_STRIDE : nullptr;
}

template <bool UC>
__device__ __forceinline__ void metric_set(const double2 a, const double2 t, double& ra, double& rb) {
  if (UC) {
    const double dx = t.x - a.x, dy = t.y - a.y;
    ra = dx * dx + dy * dy; rb = 0.0;
  } else {
    ra = a.x * t.x + a.y * t.y;
    rb = a.x * t.y - a.y * t.x;
  }
}

// Phase 3 of every kernel: finalize (reward shaping), per-env scalar outputs, Monitor latch,
// auto-reset of the control registers, and the fused K3 statistics reduction.
// Must be called by ALL threads of the CTA (the statistics use full-warp collectives).
template <bool UC, int E>
__device__ __forceinline__ void phase3(const qcd_batch_t& b, const qcd_tape_out_t& out, EnvCtl& e,
                                     const bool ctl, const long cenv, const int t, const int T,
                                     const long N, const int tid, const double2 r, const bool have_metric,
                                     double* __restrict__ stats) {
  const bool autoreset = b.flags & QCD_F_AUTORESET;
  int st_fin = 0, st_ok = 0, st_bad = 0, st_trunc = 0, st_l = 0, st_d = 0, st_o = 0, st_q = 0;
  double st_r = 0.0, st_m = 0.0, st_c = 0.0;
  if (ctl) {
    const StepOut o = finalize_step<UC>(b, e, r.x, r.y, have_metric);
    const bool done = (e.status == QCD_ST_OK) && (e.term || e.trunc);
    if (out.reward) out.reward[(long)t * N + cenv] = o.reward;
    if (out.metric) out.metric[(long)t * N + cenv] = o.metric;
    if (out.cost) out.cost[(long)t * N + cenv] = o.cost;
    if (out.depth) out.depth[(long)t * N + cenv] = e.depth;
    if (out.terminated) out.terminated[(long)t * N + cenv] = e.term;
    if (out.truncated) out.truncated[(long)t * N + cenv] = e.trunc;
    if (out.status) out.status[(long)t * N + cenv] = (int8_t)e.status;
    if (out.operations) out.operations[(long)t * N + cenv] = e.ops;
    if (out.used_wires) out.used_wires[(long)t * N + cenv] = e.used_cnt;
    if (t == T - 1) {
      b.reward[cenv] = o.reward; b.metric[cenv] = o.metric; b.cost[cenv] = o.cost;
      b.depth[cenv] = e.depth; b.operations[cenv] = e.ops; b.used_wires[cenv] = e.used_cnt;
      b.terminated[cenv] = e.term; b.truncated[cenv] = e.trunc; b.status[cenv] = (int8_t)e.status;
    }
    st_ok = e.status == QCD_ST_OK; st_bad = !st_ok;
    if (done) {
      st_fin = 1; st_trunc = e.trunc; st_l = e.ep_len; st_d = e.depth; st_o = e.ops; st_q = e.used_cnt;
      st_r = e.ep_ret; st_m = o.metric; st_c = o.cost;
      if (b.episode_return) { b.episode_return[cenv] = e.ep_ret; b.episode_length[cenv] = e.ep_len; }
      if (autoreset) {                          // env.py:117-121 fused
        e.meta = make_uint4(0, 0, 0, 0); e.last_m = 0; e.last_c = 0; e.ep_ret = 0; e.ep_len = 0;
      }
    }
  }
  // Fused K3: Monitor episode fields + termination histogram summed over the warp, then one
  // atomicAdd per non-zero field per warp of control threads (warp-uniform branch).
  if (stats != nullptr && (tid >> 5) <= ((E - 1) >> 5)) {
    constexpr unsigned FULL = 0xffffffffu;
    const int n_fin = __reduce_add_sync(FULL, st_fin);
    const int n_ok = __reduce_add_sync(FULL, st_ok);
    const int n_bad = __reduce_add_sync(FULL, st_bad);
    if (n_fin) {
      const int n_trunc = __reduce_add_sync(FULL, st_trunc);
      const int s_l = __reduce_add_sync(FULL, st_l), s_d = __reduce_add_sync(FULL, st_d);
      const int s_o = __reduce_add_sync(FULL, st_o), s_q = __reduce_add_sync(FULL, st_q);
#pragma unroll
      for (int off = 16; off >= 1; off >>= 1) {
        st_r += __shfl_xor_sync(FULL, st_r, off);
        st_m += __shfl_xor_sync(FULL, st_m, off);
        st_c += __shfl_xor_sync(FULL, st_c, off);
      }
      if ((tid & 31) == 0) {
        atomicAdd(stats + QCD_STAT_EPISODES, (double)n_fin);
        if (b.ep_return) { atomicAdd(stats + QCD_STAT_SUM_R, st_r); atomicAdd(stats + QCD_STAT_SUM_L, (double)s_l); }
        atomicAdd(stats + QCD_STAT_SUM_D, (double)s_d);
        atomicAdd(stats + QCD_STAT_SUM_O, (double)s_o);
        atomicAdd(stats + QCD_STAT_SUM_Q, (double)s_q);
        atomicAdd(stats + QCD_STAT_SUM_M, st_m);
        atomicAdd(stats + QCD_STAT_SUM_C, st_c);
        if (n_fin - n_trunc) atomicAdd(stats + QCD_STAT_N_DONE, (double)(n_fin - n_trunc));
        if (n_trunc) atomicAdd(stats + QCD_STAT_N_DEPTH, (double)n_trunc);
      }
    }
    if ((tid & 31) == 0) {
      if (n_ok) atomicAdd(stats + QCD_STAT_STEPS, (double)n_ok);
      if (n_bad) atomicAdd(stats + QCD_STAT_BAD_ACTIONS, (double)n_bad);
    }
  }
}

template <int LOG2S, bool UC, bool STAGED, int LOG2L, int NT>
struct KCfg {
  static constexpr int S = 1 << LOG2S;
  static constexpr int L = 1 << LOG2L;            // threads per env
  static constexpr int E = NT / L;                // envs per CTA
  static constexpr int PPT = (S / 2) / L;         // amplitude pairs per thread
  static constexpr int U = PPT < 4 ? PPT : 4;     // pairs in flight per thread
  static_assert(PPT >= 1, "group larger than the number of pairs");
  static_assert(L <= 32 || E == 1, "a group wider than a warp owns the whole CTA");
};

// ------------------------------------------------------------------------------------------------
template <int LOG2S, bool UC, bool STAGED, int LOG2L, int NT>
__global__ void __launch_bounds__(NT)
qcd_kernel(const qcd_batch_t b, const void* __restrict__ actions, const int act_f32, const int T,
           const qcd_tape_out_t out, const int metric_every_step, const double4 imm) {
  using C = KCfg<LOG2S, UC, STAGED, LOG2L, NT>;
  constexpr int S = C::S, L = C::L, E = C::E, PPT = C::PPT, U = C::U;

  __shared__ GateRec s_rec[E];
  __shared__ double2 s_red[E];
  __shared__ double2 s_wred[NT / 32];
  extern __shared__ double2 s_tile[];             // STAGED: [E][S]

  const int tid = threadIdx.x;
  const long env0 = (long)blockIdx.x * E;
  const long N = b.n_envs;
  const int grp = tid >> LOG2L, lane = tid & (L - 1);
  const long genv = env0 + grp;                   // env of this thread's group
  const bool gact = genv < N;
  const bool ctl = tid < E && env0 + tid < N;     // this thread is the control thread of env0+tid
  const long cenv = env0 + tid;
  const bool per_env_tgt = b.flags & QCD_F_TARGET_PER_ENV;

  double2* gstate = reinterpret_cast<double2*>(b.state);
  const double2* tgt = reinterpret_cast<const double2*>(b.target) + (per_env_tgt && gact ? genv * S : 0);
  double2* st = STAGED ? (s_tile + (long)grp * S) : (gstate + (gact ? genv * S : 0));

  // ---- control-thread persistent registers
  EnvCtl e;
  e.meta = make_uint4(0, 0, 0, 0); e.last_m = 0; e.last_c = 0; e.ep_ret = 0; e.ep_len = 0;
  if (ctl) {
    e.meta = reinterpret_cast<const uint4*>(b.meta)[cenv];
    if (b.last) { const double2 l2 = reinterpret_cast<const double2*>(b.last)[cenv]; e.last_m = l2.x; e.last_c = l2.y; }
    if (b.ep_return) { e.ep_ret = b.ep_return[cenv]; e.ep_len = b.ep_length[cenv]; }
  }

  if (!STAGED && QCD_PAIR_L2_PREFETCH && S <= QCD_PAIR_L2_PREFETCH_MAX_S) {
    // The pair addresses depend on the decoded wire, so the state loads cannot be issued before the
    // barrier below; pull the env's lines into L2 meanwhile (128-byte lines, L lanes per env).
    if (gact) {
      const char* base = reinterpret_cast<const char*>(gstate + genv * S);
#pragma unroll 1
      for (int ln = lane; ln < S * 16 / 128; ln += L)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(base + (long)ln * 128));
    }
  }
  if (STAGED) {                                   // stage the CTA's contiguous block of states
    const long base = env0 * S;
    const long lim = (N - env0 < E ? N - env0 : E) * (long)S;
    for (long i = tid; i < lim; i += NT) s_tile[i] = gstate[base + i];
  }

  for (int t = 0; t < T; ++t) {
    // ================= phase 1: per-env scalar decode
    if (ctl) {
      double a0, a1, a2, a3;
      if (act_f32 == ACT_IMMEDIATE) { a0 = imm.x; a1 = imm.y; a2 = imm.z; a3 = imm.w; }   // single env: action by value
      else load_action(actions, act_f32, (long)t * N + cenv, a0, a1, a2, a3);
      GateRec rec;
      decode_action(b, a0, a1, a2, a3, e, rec);
      const bool done = e.term || e.trunc;
      if (metric_every_step || done || t == T - 1) rec.flags |= RF_METRIC;
      if (t == T - 1) rec.flags |= RF_LAST;
      s_rec[tid] = rec;
    }
    __syncthreads();

    // ================= phase 2: the group's butterflies, metric partial sums, stores
    double ra = 0.0, rb = 0.0;
    if (gact) {
      const GateRec g = s_rec[grp];
      const bool reset = g.flags & RF_RESET;
      const bool want_metric = g.flags & RF_METRIC;
      const bool last = g.flags & RF_LAST;
      const bool wr_state = STAGED ? true : (g.kind != K_NONE || reset);
      float* ob = (b.obs && last) ? b.obs + genv * b.obs_stride : nullptr;
      float* fob = (b.final_obs && reset) ? b.final_obs + genv * b.obs_stride : nullptr;   // every reset of a tape
      const unsigned lowmask = (1u << g.wbit) - 1u;
#pragma unroll 1
      for (int j0 = 0; j0 < PPT; j0 += U) {
        int i0[U];
        double2 x0[U], x1[U], t0[U], t1[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const unsigned k = (unsigned)(lane + L * (j0 + u));
          i0[u] = (int)(((k & ~lowmask) << 1) | (k & lowmask));
          x0[u] = st[i0[u]];
          x1[u] = st[i0[u] | (1 << g.wbit)];
        }
        if (want_metric) {
#pragma unroll
          for (int u = 0; u < U; ++u) {
            t0[u] = __ldg(tgt + i0[u]);
            t1[u] = __ldg(tgt + (i0[u] | (1 << g.wbit)));
          }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int i1 = i0[u] | (1 << g.wbit);
          apply_pair(g, i0[u], x0[u], x1[u]);
          if (want_metric) {
            metric_acc<UC>(x0[u], t0[u], ra, rb);
            metric_acc<UC>(x1[u], t1[u], ra, rb);
          }
          if (fob) {
            fob[i0[u]] = (float)x0[u].x; fob[S + i0[u]] = (float)x0[u].y;
            fob[i1] = (float)x1[u].x;    fob[S + i1] = (float)x1[u].y;
          }
          if (reset) {
            x0[u] = reset_amp(UC, b.n_qubits, i0[u]);
            x1[u] = reset_amp(UC, b.n_qubits, i1);
          }
          if (wr_state) {
            // store only what the gate changed: P/CP leave a0 (and a1 without the control bit)
            // untouched, CX only swaps pairs whose control bit is set.  Whole 32-byte sectors stay
            // clean for every wire bit >= 1.
            const bool cset = (i0[u] >> g.cbit) & 1;
            const bool ch1 = STAGED || reset || g.kind == K_RX || g.kind == K_P || ((g.kind == K_CP || g.kind == K_CX) && cset);
            const bool ch0 = STAGED || reset || g.kind == K_RX || (g.kind == K_CX && cset);
            if (ch0) st[i0[u]] = x0[u];
            if (ch1) st[i1] = x1[u];
          }
          if (ob) {
            ob[i0[u]] = (float)x0[u].x; ob[S + i0[u]] = (float)x0[u].y;
            ob[i1] = (float)x1[u].x;    ob[S + i1] = (float)x1[u].y;
          }
        }
      }
    }
    // group reduction of (ra, rb): all lanes of every warp take part
    if (L > 1) {
#pragma unroll
      for (int off = (L < 32 ? L : 32) / 2; off >= 1; off >>= 1) {
        ra += __shfl_xor_sync(0xffffffffu, ra, off);
        rb += __shfl_xor_sync(0xffffffffu, rb, off);
      }
    }
    if (L > 32) {
      if ((tid & 31) == 0) s_wred[tid >> 5] = make_double2(ra, rb);
    } else if (lane == 0 && gact) {
      s_red[grp] = make_double2(ra, rb);
    }
    __syncthreads();
    if (L > 32 && tid == 0) {
      double xa = 0.0, xb = 0.0;
      for (int wv = 0; wv < NT / 32; ++wv) { xa += s_wred[wv].x; xb += s_wred[wv].y; }
      s_red[0] = make_double2(xa, xb);            // read below by the same thread (tid 0 is ctl)
    }

    // ================= phase 3: per-env finalize + scalar outputs + fused episode statistics
    phase3<UC, E>(b, out, e, ctl, cenv, t, T, N, tid, ctl ? s_red[tid] : make_double2(0.0, 0.0),
                  ctl && (s_rec[tid].flags & RF_METRIC), stats_slice(b, blockIdx.x));
    // No barrier needed here: s_rec[tid] / s_red[tid] are re-written in phase 1 / phase 2 of step
    // t+1 only after their phase-2 / phase-3 readers of step t have passed the barrier above /
    // are the writing thread itself.
  }

  if (ctl) {
    reinterpret_cast<uint4*>(b.meta)[cenv] = e.meta;
    if (b.last) reinterpret_cast<double2*>(b.last)[cenv] = make_double2(e.last_m, e.last_c);
    if (b.ep_return) { b.ep_return[cenv] = e.ep_ret; b.ep_length[cenv] = e.ep_len; }
  }
  if (STAGED) {
    const long base = env0 * S;
    const long lim = (N - env0 < E ? N - env0 : E) * (long)S;
    for (long i = tid; i < lim; i += NT) gstate[base + i] = s_tile[i];
  }
}

// ------------------------------------------------------------------------------------------------
// K2b, round 2: replay for S >= 512 with each env's state resident in REGISTERS for the whole tape.
//
// ncu on the round-1 kernel (state in shared memory, one sweep and one barrier per gate; profiles/r01_summary.md §4): 4 695 warp-instructions per env-step = 73 per
// amplitude pair where an RX pair needs ~14; the index arithmetic, the per-pair switch, the shared-memory
// round trip of every sweep and the bookkeeping replicated in 256 threads were the cost, not FP64.  Here
//   * a CTA of S/16 threads owns one env; every thread keeps 16 amplitudes (4 "local" index bits) in
//     registers.  The map position -> index bit is fixed at compile time (RegMap): positions 0..3 are the
//     register bits, 4..8 the lane bits, 9.. the warp bits.  For UC the six (eta) ROW bits — the only bits
//     a gate can act on — come first, so every UC gate is either pure register work (4 of 6 wires at
//     eta=6) or one shuffle per amplitude (2 of 6); shared memory is not touched by a gate at all.  For SP
//     the lane bits are index bits 0..4 (coalesced I/O), the register bits 5..8, and a gate on a warp bit
//     (3 of 12 at eta=12) exchanges amplitudes through shared memory — diagonal gates never move data;
//   * the gate of a step is dispatched ONCE per step, through an opcode the decode-ahead resolved (ROP_*), to code
//     specialised on (register bit, gate kind; both register bits of a register-register CP; the control's register
//     bit of a CX on a lane bit): no per-pair switch, no index arithmetic (the 16 register slots are compile-time);
//     steps without a gate skip the dispatch;
//   * the depth stack lives in the lanes (lane q = qubit q): two shuffles and a max per gate, the circuit depth is the
//     scalar recurrence max(depth, new level): every warp has the depth / truncation / reset decision without a
//     broadcast or barrier;
//   * the action-only part of the decode (floor, kind, sincos) is done one step per thread ahead of time;
//   * only steps that need the metric (episode end, last step; every step when dense or recorded) evaluate it; they
//     leave one partial sum per warp in shared memory and are finalized per chunk, one metric step per thread
//     (reg_finalize_chunk: atan, cost, shaping, records, statistics); only a dense reward / a carried return needs
//     the short sequential pass of thread 0;
//   * UC: every thread parks its 16 target amplitudes in shared memory; statistics are flushed once per launch.
// Bit-identical amplitudes to qcd_step (same mul_phase / rot_rx).
template <int LOG2S, bool UC>
struct RegMap {
  static constexpr int S = 1 << LOG2S;
  static constexpr int NL = 4, A = 1 << NL;            // register bits, amplitudes per thread
  static constexpr int NT = S >> NL;                   // threads per CTA (one env)
  static constexpr int NW = NT / 32;                   // warps
  static constexpr int CH = 128;                       // steps decoded (and finalized) per chunk
  static constexpr int ETA = UC ? LOG2S / 2 : LOG2S;
  static_assert(NT >= 32, "replay_reg_kernel needs S >= 512");
  __host__ __device__ static constexpr int bit_of_pos(const int p) {
    return UC ? (p < ETA ? ETA + p : p - ETA)          // row bits first, then the column bits
              : (p < 4 ? 5 + p : (p < 9 ? p - 4 : p)); // SP: registers 5..8, lanes 0..4, warps 9..
  }
  __host__ __device__ static constexpr int pos_of_bit(const int b) {
    for (int p = 0; p < LOG2S; ++p) if (bit_of_pos(p) == b) return p;
    return 0;
  }
  __host__ __device__ static constexpr unsigned long long pos_table() {   // 4 bits per index bit
    unsigned long long t = 0;
    for (int b = 0; b < LOG2S; ++b) t |= (unsigned long long)pos_of_bit(b) << (4 * b);
    return t;
  }
  __host__ __device__ static constexpr int koff(const int k) {            // index offset of register slot k
    int o = 0;
    for (int j = 0; j < NL; ++j) if ((k >> j) & 1) o |= 1 << bit_of_pos(j);
    return o;
  }
  // a gate can only target a warp position when one of the ACTIVE bits sits there
  static constexpr bool WARP_TARGETS = UC ? (ETA > 9) : (LOG2S > 9);
  static constexpr size_t XCHG_BYTES = WARP_TARGETS ? (size_t)S * sizeof(double2) : 0;
  // UC: a warp's 16-byte target loads of one slot touch four separate 128-byte lines (two lane bits are the high
  // row bits), which costs the L1 data pipe 4x the cycles of a contiguous access; every thread therefore parks ITS
  // 16 target amplitudes in shared memory once per launch ([slot][thread]: conflict-free, private, no barrier).
  // SP lanes are index bits 0..4: the loads are contiguous and stay on the L1 path.
  static constexpr bool TGT_SMEM = UC;
  static constexpr size_t TGT_BYTES = TGT_SMEM ? (size_t)S * sizeof(double2) : 0;
  static constexpr size_t DYN_BYTES = XCHG_BYTES + TGT_BYTES;
};

struct __align__(16) RegStep { double cr, sr; int aux; int idx; int pk; int pad; };
// Everything the step loop needs from an action, resolved by the decode-ahead so that the loop extracts no bit fields
// it does not use:
// pk:  [2:0] kind [4:3] status [5] term [22] metric needed whatever happens [27:23] opcode of the gate (ROP_*: ONE
//      dispatch per step)
// aux: [15:0] the 16 register slots the CONTROL allows (all of them unless the control is a register bit: then the slots
//      with that bit set) [27:16] wire mask (1 << w | 1 << c)
// idx: byte 0 = w, byte 1 = c (shuffle source lanes of the depth stack: the hardware takes them modulo 32), byte 2 = xor
//      mask of the partner THREAD for a target on a lane / warp bit, byte 3 = the thread-index bits (tid) that must be set
//      for the gate to act in this thread at all (thread-bit control; thread-bit target of a diagonal gate)
//
// Opcodes, resolved by the decode-ahead (they depend on the action alone): the target's register bit J and the
// gate kind are compile-time in every case, and a CP between two register bits knows its four slots.
constexpr int ROP_NONE = 0;
constexpr int ROP_P = 1;        // + J: P on register bit J; also CP whose control is a thread bit (thread-predicated)
constexpr int ROP_RX = 5;       // + J
constexpr int ROP_CX = 9;       // + J: CX with the target on register bit J (control: register or thread bit)
constexpr int ROP_CP2 = 13;     // + pair index of (J < J'): CP between two register bits
constexpr int ROP_DIAG_T = 19;  // P / CP whose qubits are all thread bits: all 16 slots or none
constexpr int ROP_LANE = 20;    // RX / CX with the target on a lane bit: one shuffle per amplitude
constexpr int ROP_WARP = 21;    // RX / CX with the target on a warp bit (SP): exchange through shared memory
constexpr int ROP_LANE_CX = 22; // + J': CX with the target on a lane bit and the control on register bit J': only
                                // the eight slots with that bit set are exchanged
__host__ __device__ constexpr int cp2_index(const int j, const int j2) {   // j < j2 < 4 -> 0..5
  return j == 0 ? j2 - 1 : (j == 1 ? j2 + 1 : 5);
}

// slots k (of 16) whose bit j is set: 0xAAAA, 0xCCCC, 0xF0F0, 0xFF00
__device__ __forceinline__ unsigned local_mask(const int j) { return (unsigned)(0xFF00F0F0CCCCAAAAull >> (16 * j)) & 0xffffu; }

template <int J>
__device__ __forceinline__ void reg_gate_p(const double cr, const double sr, double2 (&a)[16]) {
#pragma unroll
  for (int k = 0; k < 16; ++k) if (k & (1 << J)) a[k] = mul_phase(a[k], cr, sr);
}
template <int J, int J2>
__device__ __forceinline__ void reg_gate_cp2(const double cr, const double sr, double2 (&a)[16]) {
#pragma unroll
  for (int k = 0; k < 16; ++k) if ((k & (1 << J)) && (k & (1 << J2))) a[k] = mul_phase(a[k], cr, sr);
}
template <int J>
__device__ __forceinline__ void reg_gate_rx(const double cr, const double sr, double2 (&a)[16]) {
  constexpr int B = 1 << J;
#pragma unroll
  for (int k = 0; k < 16; ++k) if (!(k & B)) {
    // rot_rx on both members of the pair, written so that every result can take its own operand's register: the four
    // products first, then each FMA overwrites the component it read (same operations, same rounding as rot_rx)
    double2& lo = a[k];
    double2& hi = a[k | B];
    const double t1 = __dmul_rn(sr, hi.y), t2 = __dmul_rn(sr, hi.x), t3 = __dmul_rn(sr, lo.y), t4 = __dmul_rn(sr, lo.x);
    lo.x = __fma_rn(cr, lo.x, t1); lo.y = __fma_rn(cr, lo.y, -t2);
    hi.x = __fma_rn(cr, hi.x, t3); hi.y = __fma_rn(cr, hi.y, -t4);
  }
}
// swap the pair where the control bit is set (same for both members: c != w)
template <int J>
__device__ __forceinline__ void reg_gate_cx(const unsigned cmask, double2 (&a)[16]) {
  constexpr int B = 1 << J;
#pragma unroll
  for (int k = 0; k < 16; ++k) if (!(k & B)) {
    const bool sw = (cmask >> k) & 1u;
    const double2 lo = a[k], hi = a[k | B];
    a[k].x = sw ? hi.x : lo.x; a[k].y = sw ? hi.y : lo.y;
    a[k | B].x = sw ? lo.x : hi.x; a[k | B].y = sw ? lo.y : hi.y;
  }
}

// env.py:106-111 on a register-resident state: this warp's share of |<psi|target>|^2 / ||U - V||_F^2
template <int LOG2S, bool UC>
__device__ __forceinline__ double2 reg_warp_metric(const double2 (&a)[16], const double2* s_tgt, const double2* tgt) {
  using M = RegMap<LOG2S, UC>;
  double pa[4] = {0.0, 0.0, 0.0, 0.0}, pb[4] = {0.0, 0.0, 0.0, 0.0};   // four independent chains
#pragma unroll
  for (int k = 0; k < M::A; ++k)
    metric_acc<UC>(a[k], M::TGT_SMEM ? s_tgt[k * M::NT] : __ldg(tgt + M::koff(k)), pa[k & 3], pb[k & 3]);
  double ra = (pa[0] + pa[1]) + (pa[2] + pa[3]), rb = (pb[0] + pb[1]) + (pb[2] + pb[3]);
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    ra += __shfl_xor_sync(0xffffffffu, ra, off);
    if (!UC) rb += __shfl_xor_sync(0xffffffffu, rb, off);
  }
  return make_double2(ra, rb);
}

// Phase 2 of replay_reg_kernel: every metric step of a chunk is finalized here, by ALL threads of the CTA, once per
// chunk.  Written to hold few values at a time: it is compiled between the 64 state registers every thread keeps
// through the step loop and the 128-register budget, and anything wider makes the LOOP spill.
template <bool UC, int NT, int NW>
__device__ __forceinline__ void reg_finalize_chunk(const qcd_batch_t& b, const qcd_tape_out_t& out,
                                                   const double2 (*s_msum)[NW], const int4* s_minfo, double2* s_mres,
                                                   double2 (*s_fin)[2], double* s_seq, int* s_istat, const int nm,
                                                   const int t0, const int T, const long env) {
  const int tid = threadIdx.x;
  const long N = b.n_envs;
  const bool autoreset = b.flags & QCD_F_AUTORESET;

  // ================= phase 2a: raw metric (sqrt / atan) and raw cost of every metric step, one step per thread
  const double carried = s_seq[0];                             // (read before thread 0 may rewrite it in phase 2b)
  for (int i = tid; i < nm; i += NT) {
    double xa = 0.0, xb = 0.0;
#pragma unroll
    for (int wv = 0; wv < NW; ++wv) { xa += s_msum[i][wv].x; xb += s_msum[i][wv].y; }
    const int info = s_minfo[i].y;
    s_mres[i] = make_double2(metric_from_sums<UC>(xa, xb), depth_cost(b, info & 0xff));
    if (((info >> 16) & 3) == QCD_ST_OK && ((info >> 18) & 3)) atomicMax(&s_istat[7], i);   // last finished episode
  }
  __syncthreads();
  // ================= phase 2b: what depends on the EARLIER steps of the episode (env.py:102-104: deltas of a dense
  // reward; Monitor's running return), thread 0, results into s_fin.  A sparse reward is zero until the episode
  // ends (env.py:134), so with auto-reset and no return carried in every metric step stands alone: skipped.
  const bool standalone = (b.flags & QCD_F_SPARSE) && autoreset && carried == 0.0;
  if (!standalone) {
    if (tid == 0) {
      EnvCtl e;
      e.meta = make_uint4(0, 0, 0, 0);
      e.ep_ret = s_seq[0]; e.last_m = s_seq[1]; e.last_c = s_seq[2];
      for (int i = 0; i < nm; ++i) {
        const int info = s_minfo[i].y;
        e.status = (info >> 16) & 3; e.term = (info >> 18) & 1; e.trunc = (info >> 19) & 1;
        e.ep_len = 0; e.cost_raw = s_mres[i].y;
        const StepOut o = shape_step(b, e, s_mres[i].x);
        s_fin[i][0] = make_double2(o.reward, o.metric);
        s_fin[i][1] = make_double2(o.cost, e.ep_ret);
        if (e.status == QCD_ST_OK && (e.term || e.trunc) && autoreset) { e.ep_ret = 0.0; e.last_m = 0.0; e.last_c = 0.0; }
      }
      s_seq[0] = e.ep_ret; s_seq[1] = e.last_m; s_seq[2] = e.last_c;
    }
    __syncthreads();
  }
  // ================= phase 2c: shaping of stand-alone steps (env.py:133-135), records, Monitor latch, one metric step
  // per thread; a finished episode counts into the integer statistics and leaves (r, m, c) in its s_fin slot
  for (int i = tid; i < nm; i += NT) {
    const int4 mi = s_minfo[i];
    const int t = t0 + mi.x, depth = mi.y & 0xff, used_cnt = (mi.y >> 8) & 0xff, status = (mi.y >> 16) & 3;
    const bool term = (mi.y >> 18) & 1, trunc = (mi.y >> 19) & 1;
    StepOut o;
    double ep_ret;
    if (standalone) {
      EnvCtl e;
      e.meta = make_uint4(0, 0, 0, 0); e.ep_ret = 0.0; e.last_m = 0.0; e.last_c = 0.0; e.ep_len = 0;
      e.status = status; e.term = term; e.trunc = trunc; e.cost_raw = s_mres[i].y;
      o = shape_step(b, e, s_mres[i].x);
      ep_ret = e.ep_ret;
    } else {
      const double2 f0 = s_fin[i][0], f1 = s_fin[i][1];
      o.reward = f0.x; o.metric = f0.y; o.cost = f1.x; ep_ret = f1.y;
    }
    const long at = (long)t * N + env;
    if (out.reward) out.reward[at] = o.reward;
    if (out.metric) out.metric[at] = o.metric;
    if (out.cost) out.cost[at] = o.cost;
    if (out.depth) out.depth[at] = depth;
    if (out.terminated) out.terminated[at] = term;
    if (out.truncated) out.truncated[at] = trunc;
    if (out.status) out.status[at] = (int8_t)status;
    if (out.operations) out.operations[at] = mi.z;
    if (out.used_wires) out.used_wires[at] = used_cnt;
    if (t == T - 1) {
      b.reward[env] = o.reward; b.metric[env] = o.metric; b.cost[env] = o.cost;
      b.depth[env] = depth; b.operations[env] = mi.z; b.used_wires[env] = used_cnt;
      b.terminated[env] = term; b.truncated[env] = trunc; b.status[env] = (int8_t)status;
    }
    const bool done = status == QCD_ST_OK && (term || trunc);
    if (done) {
      const int ep_l = mi.w + 1;                               // (recorded before this step was counted, monitor.py:38)
      // Monitor 'r' / 'l' of the env: the chunk's LAST finished episode stays (later chunks overwrite)
      if (b.episode_return && i == s_istat[7]) { b.episode_return[env] = ep_ret; b.episode_length[env] = ep_l; }
      atomicAdd(&s_istat[0], 1); atomicAdd(&s_istat[1], ep_l); atomicAdd(&s_istat[2], depth);
      atomicAdd(&s_istat[3], mi.z); atomicAdd(&s_istat[4], used_cnt); atomicAdd(&s_istat[trunc ? 6 : 5], 1);
    }
    s_fin[i][0] = make_double2(done ? ep_ret : 0.0, done ? o.metric : 0.0);
    s_fin[i][1] = make_double2(done ? o.cost : 0.0, 0.0);
  }
  __syncthreads();
  // ================= phase 2d: the float64 statistics, summed in step order (three threads, one field each)
  if (tid < 3) {
    const double* f = reinterpret_cast<const double*>(&s_fin[0][0]) + tid;
    double acc = s_seq[4 + (tid == 0 ? QCD_STAT_SUM_R : (tid == 1 ? QCD_STAT_SUM_M : QCD_STAT_SUM_C))];
    for (int i = 0; i < nm; ++i) acc += f[4 * i];
    s_seq[4 + (tid == 0 ? QCD_STAT_SUM_R : (tid == 1 ? QCD_STAT_SUM_M : QCD_STAT_SUM_C))] = acc;
  }
}

template <int LOG2S, bool UC, int MINB>
__global__ void __launch_bounds__(RegMap<LOG2S, UC>::NT, MINB)
replay_reg_kernel(const qcd_batch_t b, const void* __restrict__ actions, const int act_f32, const int T,
                  const qcd_tape_out_t out, const int metric_every_step) {
  using M = RegMap<LOG2S, UC>;
  constexpr int S = M::S, NT = M::NT, NW = M::NW, A = M::A, CH = M::CH;
  constexpr unsigned FULL = 0xffffffffu;
  constexpr unsigned long long POS = M::pos_table();
  extern __shared__ __align__(16) double2 s_dynamic[];
  double2* const s_xchg = s_dynamic;                          // [S] exchange buffer (SP gates on warp bits)
  double2* const s_tgt = s_dynamic + M::XCHG_BYTES / sizeof(double2) + threadIdx.x;   // [16][NT] this thread's target slots
  __shared__ RegStep s_tab[CH + 1];                           // decoded steps of the chunk (+1: the loop fetches one ahead)
  __shared__ double2 s_msum[CH][NW];                          // per-warp metric sums of the chunk's metric steps
  __shared__ int4 s_minfo[CH];                                // their bookkeeping: {t, depth|used<<8|flags<<16, ops, ep_len}
  __shared__ double2 s_mres[CH];                              // their (raw metric, raw cost)
  __shared__ double2 s_fin[CH][2];                            // their (reward, metric), (cost, ep_return) when sequential

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long env = blockIdx.x;
  const long N = b.n_envs;
  const bool autoreset = b.flags & QCD_F_AUTORESET;
  const int shift = UC ? b.n_qubits : 0;
  const unsigned tpat = (unsigned)tid << 4;                   // bit p (p >= 4) = this thread's bit at position p
  int tbase = 0;
#pragma unroll
  for (int p = 4; p < LOG2S; ++p) tbase |= (int)((tpat >> p) & 1u) << M::bit_of_pos(p);
  double2* gstate = reinterpret_cast<double2*>(b.state) + env * S + tbase;
  const double2* tgt = reinterpret_cast<const double2*>(b.target) +
                       ((b.flags & QCD_F_TARGET_PER_ENV) ? env * S : 0) + tbase;

  if (M::TGT_SMEM) {
#pragma unroll
    for (int k = 0; k < A; ++k) s_tgt[k * NT] = __ldg(tgt + M::koff(k));
  }
  unsigned rmask = 0;                                         // slots that hold 1 after a reset (e0 / I)
#pragma unroll
  for (int k = 0; k < A; ++k) rmask |= (reset_amp(UC, M::ETA, tbase + M::koff(k)).x != 0.0 ? 1u : 0u) << k;
  double2 a[A];
  // the metric sums of the RESET state, once per launch: with the reference's own action sampler a third of the
  // episodes ends on its first step (Terminate), and those are evaluated on e0 / I
  __shared__ double2 s_m0[NW];
  {
#pragma unroll
    for (int k = 0; k < A; ++k) a[k] = make_double2(((rmask >> k) & 1u) ? 1.0 : 0.0, 0.0);
    const double2 r0 = reg_warp_metric<LOG2S, UC>(a, s_tgt, tgt);
    if (lane == 0) s_m0[warp] = r0;
    __syncwarp();
  }
  bool fresh = false;                                         // the registers hold exactly the reset state
#pragma unroll
  for (int k = 0; k < A; ++k) a[k] = gstate[M::koff(k)];

  // bookkeeping, replicated per WARP: lane q holds the depth of qubit q (QuantumCircuit.depth stack)
  int d = lane < 12 ? (int)reinterpret_cast<const uint8_t*>(b.meta)[env * 16 + lane] : 0;
  const uint32_t mw = reinterpret_cast<const uint32_t*>(b.meta)[env * 4 + 3];
  uint32_t ops = mw & 0xffffu, used = mw >> 16;
  int ep_len = b.ep_return ? b.ep_length[env] : 0;
  int depth = __reduce_max_sync(FULL, d);                     // circuit depth = max over the qubits' counters
  // sequential episode state (thread 0, phase 2b) + statistics of the launch (phase 2c): shared memory
  __shared__ double s_seq[4 + QCD_STATS_LEN];                 // ep_return, last_metric, last_cost, -, statistics
  __shared__ int s_nbad;                                      // steps with an invalid action (counted by the decode)
  // integer statistics of the launch {episodes, sum_l, sum_d, sum_o, sum_q, n_done, n_depth}; [7]: metric-step index
  // of the current chunk's last finished episode
  __shared__ int s_istat[8];
  if (tid == 0) {
    s_nbad = 0;
    s_seq[0] = b.ep_return ? b.ep_return[env] : 0.0;
    const double2 l2 = b.last ? reinterpret_cast<const double2*>(b.last)[env] : make_double2(0.0, 0.0);
    s_seq[1] = l2.x; s_seq[2] = l2.y;
#pragma unroll
    for (int k = 0; k < QCD_STATS_LEN; ++k) s_seq[4 + k] = 0.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) s_istat[k] = 0;
  }
  __syncthreads();                                            // s_nbad is zero before the decode below counts into it

  for (int t0 = 0; t0 < T; t0 += CH) {
    // ---- decode-ahead: step t0 + tid (env.py:90-100, the part that depends on the action alone)
    const int nsteps = (T - t0) < CH ? (T - t0) : CH;
    for (int i = tid; i < nsteps; i += NT) {
      double a0, a1, a2, a3;
      load_action(actions, act_f32, (long)(t0 + i) * N + env, a0, a1, a2, a3);
      const GateKind k = decode_kind(b.n_qubits, a0, a1, a2, a3);
      RegStep r;
      r.cr = k.cr; r.sr = k.sr;
      int tpos = (int)((POS >> (4 * (k.w + shift))) & 15ull), cpos = (int)((POS >> (4 * (k.c + shift))) & 15ull);
      // CP is symmetric in its two qubits: let the REGISTER-resident one play the target
      if (k.kind == K_CP && tpos >= 4 && cpos < 4) { const int x = tpos; tpos = cpos; cpos = x; }
      const bool controlled = k.kind == K_CP || k.kind == K_CX;
      r.aux = (int)((controlled && cpos < 4) ? local_mask(cpos) : 0xffffu) | (((1 << k.w) | (1 << k.c)) << 16);
      // the ONE dispatch of the step loop: target register bit and gate kind (and a register-register CP's slot
      // set) become compile-time in the case bodies
      int op = ROP_NONE;
      if (k.kind != K_NONE) {
        if (tpos < 4) {
          op = k.kind == K_RX ? ROP_RX + tpos : (k.kind == K_CX ? ROP_CX + tpos : ROP_P + tpos);
          if (k.kind == K_CP && cpos < 4) op = ROP_CP2 + (tpos < cpos ? cp2_index(tpos, cpos) : cp2_index(cpos, tpos));
        } else if (k.kind == K_P || k.kind == K_CP) {
          op = ROP_DIAG_T;
        } else {
          op = tpos < 9 ? ((k.kind == K_CX && cpos < 4) ? ROP_LANE_CX + cpos : ROP_LANE) : ROP_WARP;
        }
      }
      // thread-index bits the gate needs set: a control on a thread bit; the target of a diagonal gate on a thread bit
      const int tneed = ((controlled && cpos >= 4) ? 1 << (cpos - 4) : 0) | (op == ROP_DIAG_T ? 1 << (tpos - 4) : 0);
      r.idx = k.w | (k.c << 8) | ((tpos >= 4 ? 1 << (tpos - 4) : 0) << 16) | (tneed << 24);
      // bit 22: the metric of this step is needed whatever happens (recorded / dense reward / last step)
      r.pk = k.kind | (k.status << 3) | ((k.term ? 1 : 0) << 5) |
             ((metric_every_step || t0 + i == T - 1) ? (1 << 22) : 0) | (op << 23);
      r.pad = 0;
      s_tab[i] = r;
      if (k.status != QCD_ST_OK) atomicAdd(&s_nbad, 1);
    }
    __syncthreads();                                          // table ready (and phase 2 of the previous chunk done)
    int nm = 0;                                               // metric steps of this chunk so far

    // ================= phase 1: the chunk's gates on registers; NO block-level synchronisation for UC
    int n_pk = s_tab[0].pk;                                   // packed word of the next step, fetched one step ahead
#pragma unroll 1
    for (int si = 0; si < nsteps; ++si) {
      const double cr = s_tab[si].cr, sr = s_tab[si].sr;
      const int2 ai = *reinterpret_cast<const int2*>(&s_tab[si].aux);
      const int aux = ai.x, idx = ai.y;
      const int pk = n_pk;
      n_pk = s_tab[si + 1].pk;                                 // (past the chunk's last step: never used; fetching the
                                                               // whole record ahead measured 2 % slower)
      const int kind = pk & 7;
      // ---- QuantumCircuit.depth / count_ops / idle_wires (env.py:74,77,86): issued BEFORE the gate (independent of
      // it), so that the shuffle / REDUX latency hides behind the gate's arithmetic; consumed after it
      if (kind != K_NONE) {
        const int dw = __shfl_sync(FULL, d, idx), dc = __shfl_sync(FULL, d, idx >> 8);   // (source lane modulo 32)
        const int lvl = min(max(dw, dc) + 1, 255);             // saturate (only past truncation)
        if (((unsigned)aux >> lane) & 0x10000u) d = lvl;       // lane == w || lane == c
        depth = max(depth, lvl);                               // the new level is the only counter that grew
        fresh = false;
        ops = min(ops + 1u, 0xffffu);
        used |= (unsigned)aux >> 16;
      }

      if (kind != K_NONE) {                                    // (Terminate / invalid action: no gate, no dispatch)
        // does the gate act in this thread at all (thread-bit control / thread-bit target of a diagonal gate)?
        const unsigned tneed = (unsigned)idx >> 24;
        const bool pass = (~(unsigned)tid & tneed) == 0u;
        // slots (of 16) this gate touches as far as the CONTROL is concerned
        const unsigned cmask = pass ? (unsigned)aux & 0xffffu : 0u;
        const int xm = (idx >> 16) & 0xff;                     // partner thread = tid ^ xm (lane / warp targets)
        switch ((pk >> 23) & 31) {
#define QCD_ROP_J(J)                                                                        \
          case ROP_P + J: if (pass) reg_gate_p<J>(cr, sr, a); break;                           \
          case ROP_RX + J: reg_gate_rx<J>(cr, sr, a); break;                                   \
          case ROP_CX + J: reg_gate_cx<J>(cmask, a); break;
          QCD_ROP_J(0) QCD_ROP_J(1) QCD_ROP_J(2) QCD_ROP_J(3)
#undef QCD_ROP_J
          case ROP_CP2 + cp2_index(0, 1): reg_gate_cp2<0, 1>(cr, sr, a); break;
          case ROP_CP2 + cp2_index(0, 2): reg_gate_cp2<0, 2>(cr, sr, a); break;
          case ROP_CP2 + cp2_index(0, 3): reg_gate_cp2<0, 3>(cr, sr, a); break;
          case ROP_CP2 + cp2_index(1, 2): reg_gate_cp2<1, 2>(cr, sr, a); break;
          case ROP_CP2 + cp2_index(1, 3): reg_gate_cp2<1, 3>(cr, sr, a); break;
          case ROP_CP2 + cp2_index(2, 3): reg_gate_cp2<2, 3>(cr, sr, a); break;
          case ROP_DIAG_T:
            // diagonal, target (and for CP also the control) held by the thread index: all 16 slots or none
            if (pass) {
#pragma unroll
              for (int k = 0; k < A; ++k) a[k] = mul_phase(a[k], cr, sr);
            }
            break;
          case ROP_LANE: {                                     // partner = lane ^ xm: one shuffle per amplitude
#pragma unroll
            for (int k = 0; k < A; ++k) {
              double2 p;
              p.x = __shfl_xor_sync(FULL, a[k].x, xm);
              p.y = __shfl_xor_sync(FULL, a[k].y, xm);
              if (kind == K_RX) a[k] = rot_rx(a[k], p, cr, sr);
              else if (pass) a[k] = p;                         // CX whose control is a thread bit: all slots or none
            }
          } break;
#define QCD_ROP_LANE_CX(J2)                                                                 \
          case ROP_LANE_CX + J2: {                                                             \
            _Pragma("unroll") for (int k = 0; k < A; ++k) if (k & (1 << J2)) {                 \
              a[k].x = __shfl_xor_sync(FULL, a[k].x, xm);                                      \
              a[k].y = __shfl_xor_sync(FULL, a[k].y, xm);                                      \
            }                                                                                  \
          } break;
          QCD_ROP_LANE_CX(0) QCD_ROP_LANE_CX(1) QCD_ROP_LANE_CX(2) QCD_ROP_LANE_CX(3)
#undef QCD_ROP_LANE_CX
          case ROP_WARP:
            if (M::WARP_TARGETS) {                             // partner in another warp: through shared memory
              const int ptid = tid ^ xm;
#pragma unroll
              for (int k = 0; k < A; ++k) s_xchg[k * NT + tid] = a[k];
              __syncthreads();
#pragma unroll
              for (int k = 0; k < A; ++k) {
                const double2 p = s_xchg[k * NT + ptid];
                if (kind == K_RX) a[k] = rot_rx(a[k], p, cr, sr);
                else if ((cmask >> k) & 1u) a[k] = p;
              }
              __syncthreads();                                 // buffer free for the next exchange
            }
            break;
          default: break;                                      // ROP_NONE: Terminate / invalid action
        }
      }

      // ---- truncation (env.py:129), termination
      const int status = (pk >> 3) & 3;
      const bool term = (pk >> 5) & 1;
      const bool ok = status == QCD_ST_OK;
      // env.py:129 also truncates at operations >= max_depth*max_qubits*2; every layer holds at most max_qubits gates,
      // so that many operations imply depth >= 2*max_depth: the depth test alone decides
      const bool trunc = ok && depth >= b.max_depth;
      const bool done = ok && (term || trunc);
      const bool want_metric = done || ((pk >> 22) & 1);
      // the step's bookkeeping record is written by lane 0 of ONE warp, rotating over the warps (every warp holds the
      // same depth / ops / used / ep_len): warp 0 alone doing it reached the end of the chunk ~4 % after the others
      const bool scribe = lane == 0 && warp == (si & (NW - 1));

      if (want_metric) {
        // ---- env.py:106-111: partial sums of |<psi|target>|^2 / ||U - V||_F^2, one value per warp; they are
        // combined, shaped and emitted in phase 2
        // (an episode that ends right after a reset is evaluated on e0 / I: that sum is known since the prologue)
        const double2 r2 = fresh ? s_m0[warp] : reg_warp_metric<LOG2S, UC>(a, s_tgt, tgt);
        if (lane == 0) s_msum[nm][warp] = r2;
        if (scribe)
          s_minfo[nm] = make_int4(si, depth | (__popc(used) << 8) | (status << 16) | ((int)term << 18) | ((int)trunc << 19),
                                  (int)ops, ep_len);
        ++nm;
      } else if (scribe) {
        // ---- light step (sparse reward, episode goes on, nobody records the metric): reward and shaped cost
        // are 0 (env.py:134), the metric is not evaluated, ep_return is unchanged
        int tt = t0 + si;
        asm volatile("" : "+r"(tt));                           // (no strength-reduced record pointers in the loop)
        const long at = (long)tt * N + env;
        if (out.reward) out.reward[at] = 0.0;
        if (out.cost) out.cost[at] = depth_cost(b, depth);
        if (out.depth) out.depth[at] = depth;
        if (out.terminated) out.terminated[at] = term;
        if (out.truncated) out.truncated[at] = trunc;
        if (out.status) out.status[at] = (int8_t)status;
        if (out.operations) out.operations[at] = (int)ops;
        if (out.used_wires) out.used_wires[at] = __popc(used);
      }
      ep_len += ok;                                            // monitor.py:38 (every warp keeps its copy)

      if (done && autoreset) {                                 // env.py:117-121 fused
        if (b.final_obs) {                                     // terminal observation of this episode
          float* fo = b.final_obs + env * b.obs_stride + tbase;
#pragma unroll
          for (int k = 0; k < A; ++k) { fo[M::koff(k)] = (float)a[k].x; fo[S + M::koff(k)] = (float)a[k].y; }
        }
        // (the mask is made opaque here: otherwise the compiler keeps the 16 high words of the reset values in
        // registers through the whole loop, which the 128-register budget pays with spills)
        if (!fresh) {                                          // (untouched since the last reset: nothing to rewrite)
          unsigned rm = rmask;
          asm volatile("" : "+r"(rm));
#pragma unroll
          for (int k = 0; k < A; ++k)                          // high word of 1.0 where bit k is set (two instructions)
            a[k] = make_double2(__hiloint2double((int)((rm & (1u << k)) * (0x3ff00000u >> k)), 0), 0.0);
          // keeps this a real branch: if-converted into 64 selects per step, the compiler carried two copies of
          // the state through the loop (128 MOVs + spills per step)
          asm volatile("" ::: "memory");
          fresh = true;
        }
        d = 0; depth = 0; ops = 0; used = 0; ep_len = 0;
      }
    }
    if (tid == 0) s_istat[7] = -1;
    __syncthreads();                                           // every warp's sums of the chunk are in shared memory
    reg_finalize_chunk<UC, NT, NW>(b, out, s_msum, s_minfo, s_mres, s_fin, s_seq, s_istat, nm, t0, T, env);
    // (the next chunk's first barrier orders these reads of s_minfo / s_mres before their next writes)
  }

  __syncthreads();                                             // phase 2 of the last chunk (statistics) is complete
  // ---- the state (and the observation of the last step) go back to HBM once.  The addresses are rebuilt from
  // a fresh read of the CTA index, so that no pointer has to stay in a register through the step loop.
  {
    unsigned cta;
    asm volatile("mov.u32 %0, %%ctaid.x;" : "=r"(cta));
    double2* gs = reinterpret_cast<double2*>(b.state) + (long)cta * S + tbase;
    float* ob = b.obs ? b.obs + (long)cta * b.obs_stride + tbase : nullptr;
#pragma unroll
    for (int k = 0; k < A; ++k) {
      gs[M::koff(k)] = a[k];
      if (ob) { ob[M::koff(k)] = (float)a[k].x; ob[S + M::koff(k)] = (float)a[k].y; }
    }
  }
  if (warp == 0) {
    if (lane < 12) reinterpret_cast<uint8_t*>(b.meta)[env * 16 + lane] = (uint8_t)d;
    if (lane == 0) {
      reinterpret_cast<uint32_t*>(b.meta)[env * 4 + 3] = ops | (used << 16);
      if (b.ep_return) { b.ep_length[env] = ep_len; b.ep_return[env] = s_seq[0]; }
      if (b.last) reinterpret_cast<double2*>(b.last)[env] = make_double2(s_seq[1], s_seq[2]);
      if (b.stats) {
        double* stats = stats_slice(b, env);
        s_seq[4 + QCD_STAT_STEPS] = (double)(T - s_nbad); s_seq[4 + QCD_STAT_BAD_ACTIONS] = (double)s_nbad;
        s_seq[4 + QCD_STAT_EPISODES] = (double)s_istat[0]; s_seq[4 + QCD_STAT_SUM_L] = (double)s_istat[1];
        s_seq[4 + QCD_STAT_SUM_D] = (double)s_istat[2]; s_seq[4 + QCD_STAT_SUM_O] = (double)s_istat[3];
        s_seq[4 + QCD_STAT_SUM_Q] = (double)s_istat[4]; s_seq[4 + QCD_STAT_N_DONE] = (double)s_istat[5];
        s_seq[4 + QCD_STAT_N_DEPTH] = (double)s_istat[6];
        if (!b.ep_return) { s_seq[4 + QCD_STAT_SUM_R] = 0.0; s_seq[4 + QCD_STAT_SUM_L] = 0.0; }   // no Monitor accumulators
#pragma unroll
        for (int k = 0; k < QCD_STATS_LEN; ++k) if (s_seq[4 + k] != 0.0) atomicAdd(stats + k, s_seq[4 + k]);
      }
    }
  }
}

// ---- programmatic dependent launch (no-ops when the launch does not carry the attribute)
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ---- TMA (cp.async.bulk) + mbarrier helpers: 1-D bulk copies of a warp's contiguous state tile
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void bulk_load(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0, spins = 0;
  do {
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    if (!ok && ++spins > (1u << 24)) __trap();    // never hang the GPU on a lost transaction
  } while (!ok);
}
__device__ __forceinline__ void bulk_store(void* dst_gmem, const void* src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
               ::"l"(dst_gmem), "r"(smem_u32(src_smem)), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
__device__ __forceinline__ void fence_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

template <int LOG2S, int TE, bool STAGE>
struct WarpCfg {
  static constexpr int S = 1 << LOG2S;
  static constexpr int LOG2L = LOG2S == 5 ? QCD_WARP_LOG2L_S32 : (LOG2S == 6 ? QCD_WARP_LOG2L_S64 : (LOG2S < 5 ? LOG2S : 5));
  static constexpr int L = 1 << LOG2L;
  static constexpr int V = S / L, LOG2V = LOG2S - LOG2L;
  static constexpr int G = 32 / L, LOG2G = 5 - LOG2L;       // envs per round
  static constexpr int R = TE / G;                         // rounds per tile
  static constexpr int CH = STAGE ? QCD_WARP_REPLAY_CHUNK : QCD_WARP_CHUNK;
  static constexpr int C = R < CH ? R : CH;                // rounds per reduction chunk
  static constexpr int PF0 = V == 1 ? QCD_WARP_PF_V1 : (V == 2 ? QCD_WARP_PF_V2 : (V <= 4 ? QCD_WARP_PF_V4 : 1));
  static constexpr int PF = STAGE ? 1 : (PF0 < C ? PF0 : C);   // rounds of state in flight in registers
  static constexpr int ROW = 36;                           // double2 slots per scratch row
  static constexpr int SUMS = C * G, P = 32 / SUMS, VAL = L / P;
  static constexpr int TILE = STAGE ? TE * S : 0;          // double2 slots of the staged state tile
  // per-warp smem: [tile][partials C*ROW][gate (cr,sr) 32][gate word 32 ints = 8 slots][mbarrier 8 slots]
  static constexpr int WARP_SLOTS = TILE + C * ROW + 32 + 8 + 8;
  static_assert(TE >= G && TE <= 32 && R >= 1 && C % PF == 0 && R % C == 0 && V <= 8 && P >= 1 && VAL >= 1, "tile shape");
};

// gate word handed from the decode lane to the rounds:
//   [2:0] kind  [3] reset  [8:4] lane xor mask of the butterfly partner (0: partner in-lane)
//   [10:9] 1 + register bit of the wire when the partner is in-lane (0 otherwise)
//   [18:11] 1<<wbit as a mask over the amplitude index  [26:19] 1<<cbit likewise
template <int LOG2L>
__device__ __forceinline__ int pack_gate(const GateRec& g) {
  const bool in_lane = g.wbit >= LOG2L;
  const int xm = in_lane ? 0 : (1 << g.wbit);
  const int jsel = in_lane ? (1 + g.wbit - LOG2L) : 0;
  return g.kind | ((g.flags & RF_RESET) ? 8 : 0) | (xm << 4) | (jsel << 9) | ((1 << g.wbit) << 11) |
         ((1 << g.cbit) << 19);
}

// ------------------------------------------------------------------------------------------------
// Warp-autonomous STEP kernel for small states (S <= 256 amplitudes): the production step path.
//
// ncu on the two kernels above (profiles/r01_*): 30-40 % of all warp samples sit at the CTA barrier
// behind the single decode warp, issue utilisation is 24-36 %, and the launch is latency- not
// HBM-bound.  Here a WARP owns a tile of TE consecutive envs and never synchronises with anyone:
//   state tile   : STAGE: lane 0 issues ONE cp.async.bulk (TMA) of the tile's TE*S*16 contiguous
//                  bytes into shared memory, completion on an mbarrier, before anything else, so the
//                  whole tile is in flight during the scalar phase; the rounds then work in place in
//                  smem and one bulk store writes the tile back.  !STAGE: register prefetch PF rounds
//                  ahead with plain coalesced LDG.128/STG.128;
//   scalar phase : lane e < TE decodes env e of the tile (coalesced scalar I/O) and leaves the gate
//                  (one packed word + cos/sin) in the warp's smem;
//   rounds       : the tile's states are processed in R rounds of G = 32/L envs, L = min(32,S) lanes
//                  per env, V = S/L amplitudes per lane (amplitude i = li + L*j).  The butterfly
//                  partner is one shuffle away (wire bit < log2 L) or in a sibling register;
//   metric sums  : per-lane partial sums go to a per-warp smem scratch, 8 rounds at a time, and are
//                  reduced transposed (each lane adds 8 values + <=2 shuffles) instead of a 5-step
//                  butterfly per env; the sum returns to the env's decode lane;
//   finalize     : lane e shapes the reward and writes env e's scalars (coalesced), statistics are
//                  reduced with REDUX/shuffles and one atomicAdd per field per warp.
template <int LOG2S, bool UC, int WARPS, int TE, bool STAGE, int MINB>
__global__ void __launch_bounds__(WARPS * 32, MINB)
step_warp_kernel(const qcd_batch_t b, const void* __restrict__ actions, const int act_f32, const double4 imm) {
  using W = WarpCfg<LOG2S, TE, STAGE>;
  constexpr int S = W::S, LOG2L = W::LOG2L, L = W::L, V = W::V, LOG2V = W::LOG2V, G = W::G, LOG2G = W::LOG2G;
  constexpr int R = W::R, C = W::C, PF = W::PF, ROW = W::ROW, SUMS = W::SUMS, P = W::P, VAL = W::VAL;
  constexpr unsigned FULL = 0xffffffffu;

  extern __shared__ __align__(128) double2 s_dyn[];

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long tile = (long)blockIdx.x * WARPS + warp;
  const long N = b.n_envs;
  const long env_base = tile * TE;
  if (env_base >= N) return;                      // whole warp; no CTA-level synchronisation below
  const int n_valid = (int)((N - env_base) < TE ? (N - env_base) : TE);
  const int li = lane & (L - 1), gi = lane >> LOG2L;
  double2* wsm = s_dyn + (long)warp * W::WARP_SLOTS;
  double2* stile = wsm;                           // [TE*S] staged states (STAGE)
  double2* part = wsm + W::TILE;                  // [C*ROW] partial metric sums
  double2* s_cs = part + C * ROW;                 // [32] (cos, sin) of each env's gate
  int* s_pk = reinterpret_cast<int*>(s_cs + 32);  // [32] packed gate word
  uint64_t* mbar = reinterpret_cast<uint64_t*>(s_cs + 32 + 8);
  double2* tstate = reinterpret_cast<double2*>(b.state) + env_base * S;
  const double2* tgt = reinterpret_cast<const double2*>(b.target);   // shared target (per-env targets
                                                                     // take the pair-streaming kernel)

  // ---- scalar loads of env_base + lane (and, !STAGE, the state of the first PF rounds)
  const long env = env_base + lane;
  const bool ctl = lane < TE && env < N;
  double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
  EnvCtl e;
  e.meta = make_uint4(0, 0, 0, 0); e.last_m = 0; e.last_c = 0; e.ep_ret = 0; e.ep_len = 0;
  // Programmatic dependent launch: this grid may have been started while the previous kernel of the
  // stream (the previous env step) was still draining.  Everything that kernel wrote (meta, states,
  // accumulators) is only touched after the wait.
  // The ACTION buffer is only read before the wait when the caller set QCD_F_ACTIONS_STABLE (it was
  // complete before the previous kernel of the stream started, e.g. a pre-generated tape); otherwise
  // a policy / sampling kernel may have written it right before this launch, and its writes are
  // only guaranteed visible after griddepcontrol.wait.
  GateKind gk;
  const bool early = b.flags & QCD_F_ACTIONS_STABLE;
  if (!early) pdl_wait();
  if (ctl) {
    if (act_f32 == ACT_IMMEDIATE) { a0 = imm.x; a1 = imm.y; a2 = imm.z; a3 = imm.w; }     // single env: action by value
    else load_action(actions, act_f32, env, a0, a1, a2, a3);
    gk = decode_kind(b.n_qubits, a0, a1, a2, a3);
  }
  if (early) pdl_wait();
  if (STAGE) {                                    // the tile goes in flight first
    if (lane == 0) {
      mbar_init(mbar, 1);
      bulk_load(stile, tstate, (uint32_t)(n_valid * S * sizeof(double2)), mbar);
    }
  }
  if (ctl) {
    e.meta = reinterpret_cast<const uint4*>(b.meta)[env];
    if (b.last) { const double2 l2 = reinterpret_cast<const double2*>(b.last)[env]; e.last_m = l2.x; e.last_c = l2.y; }
    if (b.ep_return) { e.ep_ret = b.ep_return[env]; e.ep_len = b.ep_length[env]; }
  }
  // per-lane pointers of round 0; round r is a constant G*S (state) / G*stride (obs) further on
  const long gstride = (long)G * b.obs_stride;
  double2* sp = tstate + gi * S + li;
  const double2* stp = stile + gi * S + li;
  float* op = b.obs ? b.obs + (env_base + gi) * b.obs_stride + li : nullptr;
  float* fp = b.final_obs ? b.final_obs + (env_base + gi) * b.obs_stride + li : nullptr;

  double2 buf[PF][V];
  if (!STAGE) {
#pragma unroll
    for (int u = 0; u < PF; ++u) {
      const bool ok = u * G + gi < n_valid;
#pragma unroll
      for (int j = 0; j < V; ++j) buf[u][j] = ok ? sp[u * G * S + L * j] : make_double2(0.0, 0.0);
    }
  }
  double2 tv[V];
  double rv[V];                                   // reset amplitude (e0 / identity) of this lane's slots
#pragma unroll
  for (int j = 0; j < V; ++j) {
    tv[j] = __ldg(tgt + li + L * j);
    rv[j] = reset_amp(UC, b.n_qubits, li + L * j).x;
  }

  // ---- scalar phase: decode (env.py:90-100,129,112)
  {
    int pk = 0;
    double2 cs = make_double2(1.0, 0.0);
    if (ctl) {
      GateRec rec;
      book_gate(b, gk.status, gk.kind, gk.w, gk.c, gk.term, e);
      rec.cr = gk.cr; rec.sr = gk.sr; rec.kind = gk.kind;
      const int shift = UC ? b.n_qubits : 0;      // UC: gate acts on the row index
      rec.wbit = gk.w + shift; rec.cbit = gk.c + shift;
      rec.flags = ((e.term || e.trunc) && (b.flags & QCD_F_AUTORESET)) ? RF_RESET : 0;
      pk = pack_gate<LOG2L>(rec);
      cs = make_double2(rec.cr, rec.sr);
    }
    s_cs[lane] = cs; s_pk[lane] = pk;
  }
  __syncwarp();
  if (STAGE) mbar_wait(mbar, 0);                  // tile landed

  double2 mysum = make_double2(0.0, 0.0);
#if QCD_WARP_UNROLL_CHUNKS
#pragma unroll
#else
#pragma unroll 1
#endif
  for (int c0 = 0; c0 < R; c0 += C) {
#pragma unroll
    for (int rr = 0; rr < C; rr += PF) {
      double2 cur[PF][V];
      if (STAGE) {
        const bool ok = (c0 + rr) * G + gi < n_valid;
#pragma unroll
        for (int j = 0; j < V; ++j) cur[0][j] = ok ? stp[rr * G * S + L * j] : make_double2(0.0, 0.0);
      } else {
#pragma unroll
        for (int u = 0; u < PF; ++u)
#pragma unroll
          for (int j = 0; j < V; ++j) cur[u][j] = buf[u][j];
        if (c0 + rr + PF < R) {                   // keep the next PF rounds in flight
#pragma unroll
          for (int u = 0; u < PF; ++u) {
            const bool ok = (c0 + rr + PF + u) * G + gi < n_valid;
#pragma unroll
            for (int j = 0; j < V; ++j)
              buf[u][j] = ok ? sp[(rr + PF + u) * G * S + L * j] : make_double2(0.0, 0.0);
          }
        }
      }
#pragma unroll
      for (int u = 0; u < PF; ++u) {
        const int k = rr + u;                     // round within the chunk (compile-time)
        const int src = (c0 + k) * G + gi;        // decode lane / tile-local env of this group
        const double2 cs = s_cs[src];
        const int pk = s_pk[src];
        if constexpr (QCD_WARP_BRANCHLESS && V <= QCD_WARP_BRANCHLESS_MAX_V) {
        // Branch-free round: both candidate results are computed and selected, the partner shuffle is
        // unconditional, stores are predicated.  One basic block per group of PF rounds, so the
        // scheduler can interleave independent rounds (the warp has few siblings to hide latency).
        const int kind = pk & 7;
        const bool reset = pk & 8;
        const unsigned wm = (pk >> 11) & 0xffu, cm = (pk >> 19) & 0xffu;
        const bool act = src < n_valid;
        const bool is_rx = kind == K_RX, is_cx = kind == K_CX;
        const unsigned pm = (kind == K_P) ? wm : ((kind == K_CP) ? (wm | cm) : 0x100u);   // 0x100: never
        const int src_lane = lane ^ ((pk >> 4) & 31);
        const int jsel = (pk >> 9) & 3;
        double ra, rb;
#pragma unroll
        for (int j = 0; j < V; ++j) {
          const unsigned i = li + L * j;
          const double2 v = cur[u][j];
          double2 p = v;
          if (LOG2V > 0) {
#pragma unroll
            for (int bit = 0; bit < LOG2V; ++bit) {
              const double2 cand = cur[u][j ^ (1 << bit)];
              const bool take = jsel == bit + 1;
              p.x = take ? cand.x : p.x; p.y = take ? cand.y : p.y;
            }
          }
          p.x = __shfl_sync(FULL, p.x, src_lane);
          p.y = __shfl_sync(FULL, p.y, src_lane);
          const bool ph = (i & pm) == pm;
          const bool cxs = is_cx && (i & cm);
          const double2 m = mul_phase(v, cs.x, cs.y);
          const double2 x = rot_rx(v, p, cs.x, cs.y);
          double2 y = v;
          y.x = ph ? m.x : y.x; y.y = ph ? m.y : y.y;
          y.x = is_rx ? x.x : y.x; y.y = is_rx ? x.y : y.y;
          y.x = cxs ? p.x : y.x; y.y = cxs ? p.y : y.y;
          if (j == 0) metric_set<UC>(y, tv[j], ra, rb); else metric_acc<UC>(y, tv[j], ra, rb);
          const bool wr = act && (ph || is_rx || cxs || reset);
          if (act && reset && fp) {
            float* f = fp + (long)(c0 + k) * gstride;
            f[L * j] = (float)y.x; f[S + L * j] = (float)y.y;
          }
          const double2 o2 = make_double2(reset ? rv[j] : y.x, reset ? 0.0 : y.y);
          if (wr) sp[k * G * S + L * j] = o2;
          if (act && op) { op[L * j] = (float)o2.x; op[S + L * j] = (float)o2.y; }
        }
        part[k * ROW + lane] = make_double2(ra, rb);
        } else {
        const int kind = pk & 7;
        const bool reset = pk & 8;
        const unsigned wm = (pk >> 11) & 0xffu, cm = (pk >> 19) & 0xffu;
        const bool act = src < n_valid;
        // butterfly partner: only RX and CX read it; the shuffle needs the whole warp
        const bool need_p = kind == K_RX || kind == K_CX;
        double2 pv[V];
        if (__any_sync(FULL, need_p)) {
          const int src_lane = lane ^ ((pk >> 4) & 31);
          const int jsel = (pk >> 9) & 3;
#pragma unroll
          for (int j = 0; j < V; ++j) {
            double2 p = cur[u][j];
            if (LOG2V > 0) {
#pragma unroll
              for (int bit = 0; bit < LOG2V; ++bit) {
                const double2 cand = cur[u][j ^ (1 << bit)];
                const bool take = jsel == bit + 1;
                p.x = take ? cand.x : p.x; p.y = take ? cand.y : p.y;
              }
            }
            pv[j].x = __shfl_sync(FULL, p.x, src_lane);
            pv[j].y = __shfl_sync(FULL, p.y, src_lane);
          }
        } else {
#pragma unroll
          for (int j = 0; j < V; ++j) pv[j] = cur[u][j];
        }
        double ra, rb;
        double2 y[V];
#pragma unroll
        for (int j = 0; j < V; ++j) {
          const unsigned i = li + L * j;
          const bool bw = i & wm, bc = i & cm;
          double2 v = cur[u][j];
          switch (kind) {
            case K_P:  if (bw) v = mul_phase(v, cs.x, cs.y); break;
            case K_CP: if (bw && bc) v = mul_phase(v, cs.x, cs.y); break;
            case K_RX: v = rot_rx(v, pv[j], cs.x, cs.y); break;
            case K_CX: if (bc) v = pv[j]; break;
            default: break;
          }
          y[j] = v;
          if (j == 0) metric_set<UC>(v, tv[j], ra, rb); else metric_acc<UC>(v, tv[j], ra, rb);
        }
        part[k * ROW + lane] = make_double2(ra, rb);
        if (act) {
          if (reset) {                            // episode over: terminal obs out, e0 / I in
            if (fp) {
              float* f = fp + (long)(c0 + k) * gstride;
#pragma unroll
              for (int j = 0; j < V; ++j) { f[L * j] = (float)y[j].x; f[S + L * j] = (float)y[j].y; }
            }
#pragma unroll
            for (int j = 0; j < V; ++j) {
              const double2 v = make_double2(rv[j], 0.0);
              if (STAGE) stile[src * S + li + L * j] = v; else sp[k * G * S + L * j] = v;
              if (op) { op[L * j] = (float)rv[j]; op[S + L * j] = 0.0f; }
            }
          } else {
#pragma unroll
            for (int j = 0; j < V; ++j) {
              // store only amplitudes the gate changed (whole sectors stay clean for wire bits >= 1)
              const unsigned i = li + L * j;
              const bool bw = i & wm, bc = i & cm;
              const bool changed = STAGE ? (kind != K_NONE)
                                         : (kind == K_RX || (kind == K_P && bw) || (kind == K_CP && bw && bc) || (kind == K_CX && bc));
              if (changed) { if (STAGE) stile[src * S + li + L * j] = y[j]; else sp[k * G * S + L * j] = y[j]; }
              if (op) { op[L * j] = (float)y[j].x; op[S + L * j] = (float)y[j].y; }
            }
          }
        }
        }
        if (op) op += gstride;
      }
    }
    __syncwarp();
    // ---- transposed reduction of the chunk: sum id s = k*G + g2 (round k of the chunk, group g2)
    {
      const int sidx = lane / P, prt = lane % P;
      const int k = sidx >> LOG2G, g2 = sidx & (G - 1);
      const double2* row = part + k * ROW + g2 * L + prt;
      double xa = 0.0, xb = 0.0;
#pragma unroll
      for (int jj = 0; jj < VAL; ++jj) {
        const double2 q = row[P * ((jj + g2) & (VAL - 1))];
        xa += q.x; xb += q.y;
      }
#pragma unroll
      for (int off = P / 2; off >= 1; off >>= 1) {
        xa += __shfl_xor_sync(FULL, xa, off);
        xb += __shfl_xor_sync(FULL, xb, off);
      }
      // env (c0*G + s) of the tile is decoded by lane c0*G + s: hand the sum over
      const int want = lane - c0 * G;
      const int from = (want & (SUMS - 1)) * P;
      const double va = __shfl_sync(FULL, xa, from), vb = __shfl_sync(FULL, xb, from);
      if (want >= 0 && want < SUMS) mysum = make_double2(va, vb);
    }
    __syncwarp();
    sp += C * G * S; stp += C * G * S;
  }

  if (STAGE) {                                    // one bulk store of the whole (modified) tile
    fence_async_smem();
    __syncwarp();
    if (lane == 0) bulk_store(tstate, stile, (uint32_t)(n_valid * S * sizeof(double2)));
  }

  // ---- finalize + scalar outputs + fused statistics
  pdl_launch_dependents();                        // the next step's grid may start its scalar phase now
                                                  // (triggering at kernel entry measured 2 % slower)
  const qcd_tape_out_t none = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  phase3<UC, 32>(b, none, e, ctl, env, 0, 1, N, lane, mysum, ctl, stats_slice(b, tile));
  if (ctl) {
    reinterpret_cast<uint4*>(b.meta)[env] = e.meta;
    if (b.last) reinterpret_cast<double2*>(b.last)[env] = make_double2(e.last_m, e.last_c);
    if (b.ep_return) { b.ep_return[env] = e.ep_ret; b.ep_length[env] = e.ep_len; }
  }
  if (STAGE && lane == 0) bulk_store_wait_read(); // smem must outlive the bulk store's reads
}

// ------------------------------------------------------------------------------------------------
// K2 for small states (S <= 256): the warp-autonomous scheme with the tile RESIDENT in shared memory
// for the whole tape.  One cp.async.bulk brings the warp's TE states in, T times {lane-per-env decode,
// rounds in place in smem, transposed metric reduction, finalize + per-step records}, one bulk store
// writes the tile back, observations are produced once from the final states.  The next step's action
// is prefetched while the current step runs.  Same arithmetic as the step kernel (bit-identical states).
template <int LOG2S, bool UC, int WARPS, int TE, int MINB>
__global__ void __launch_bounds__(WARPS * 32, MINB)
replay_warp_kernel(const qcd_batch_t b, const void* __restrict__ actions, const int act_f32, const int T,
                   const qcd_tape_out_t out, const int metric_every_step) {
  using W = WarpCfg<LOG2S, TE, true>;
  constexpr int S = W::S, LOG2L = W::LOG2L, L = W::L, V = W::V, LOG2V = W::LOG2V, G = W::G, LOG2G = W::LOG2G;
  constexpr int R = W::R, C = W::C, ROW = W::ROW, SUMS = W::SUMS, P = W::P, VAL = W::VAL;
  constexpr unsigned FULL = 0xffffffffu;
  (void)metric_every_step;                        // the metric of every step is cheap here: always computed

  extern __shared__ __align__(128) double2 s_dyn[];

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long tile = (long)blockIdx.x * WARPS + warp;
  const long N = b.n_envs;
  const long env_base = tile * TE;
  if (env_base >= N) return;
  const int n_valid = (int)((N - env_base) < TE ? (N - env_base) : TE);
  const int li = lane & (L - 1), gi = lane >> LOG2L;
  double2* wsm = s_dyn + (long)warp * W::WARP_SLOTS;
  double2* stile = wsm;
  double2* part = wsm + W::TILE;
  double2* s_cs = part + C * ROW;
  int* s_pk = reinterpret_cast<int*>(s_cs + 32);
  uint64_t* mbar = reinterpret_cast<uint64_t*>(s_cs + 32 + 8);
  double2* tstate = reinterpret_cast<double2*>(b.state) + env_base * S;
  const double2* tgt = reinterpret_cast<const double2*>(b.target);

  if (lane == 0) {
    mbar_init(mbar, 1);
    bulk_load(stile, tstate, (uint32_t)(n_valid * S * sizeof(double2)), mbar);
  }

  const long env = env_base + lane;
  const bool ctl = lane < TE && env < N;
  double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
  EnvCtl e;
  e.meta = make_uint4(0, 0, 0, 0); e.last_m = 0; e.last_c = 0; e.ep_ret = 0; e.ep_len = 0;
  if (ctl) {
    load_action(actions, act_f32, env, a0, a1, a2, a3);
    e.meta = reinterpret_cast<const uint4*>(b.meta)[env];
    if (b.last) { const double2 l2 = reinterpret_cast<const double2*>(b.last)[env]; e.last_m = l2.x; e.last_c = l2.y; }
    if (b.ep_return) { e.ep_ret = b.ep_return[env]; e.ep_len = b.ep_length[env]; }
  }
  double2* stp = stile + gi * S + li;
  // terminal observations (QCD_F_AUTORESET + final_obs): written at EVERY reset of the tape, so that the
  // buffer ends up exactly as T qcd_step launches leave it
  float* fp = b.final_obs ? b.final_obs + (env_base + gi) * b.obs_stride + li : nullptr;
  const long gstride = (long)G * b.obs_stride;
  double2 tv[V];
  double rv[V];
#pragma unroll
  for (int j = 0; j < V; ++j) {
    tv[j] = __ldg(tgt + li + L * j);
    rv[j] = reset_amp(UC, b.n_qubits, li + L * j).x;
  }
  double* stats = stats_slice(b, tile);

#pragma unroll 1
  for (int t = 0; t < T; ++t) {
    // ---- scalar phase of step t; the action of step t+1 goes in flight
    {
      int pk = 0;
      double2 cs = make_double2(1.0, 0.0);
      if (ctl) {
        GateRec rec;
        decode_action(b, a0, a1, a2, a3, e, rec);
        pk = pack_gate<LOG2L>(rec);
        cs = make_double2(rec.cr, rec.sr);
        if (t + 1 < T) load_action(actions, act_f32, (long)(t + 1) * N + env, a0, a1, a2, a3);
      }
      s_cs[lane] = cs; s_pk[lane] = pk;
    }
    __syncwarp();
    if (t == 0) mbar_wait(mbar, 0);

    double2 mysum = make_double2(0.0, 0.0);
#pragma unroll 1
    for (int c0 = 0; c0 < R; c0 += C) {
#pragma unroll
      for (int k = 0; k < C; ++k) {
        const int src = (c0 + k) * G + gi;
        const double2 cs = s_cs[src];
        const int pk = s_pk[src];
        const int kind = pk & 7;
        const bool reset = pk & 8;
        const unsigned wm = (pk >> 11) & 0xffu, cm = (pk >> 19) & 0xffu;
        const bool act = src < n_valid;
        double2* sp = stp + (c0 + k) * G * S;
        double2 cur[V], pv[V];
#pragma unroll
        for (int j = 0; j < V; ++j) cur[j] = act ? sp[L * j] : make_double2(0.0, 0.0);
        const bool need_p = kind == K_RX || kind == K_CX;
        if (__any_sync(FULL, need_p)) {
          const int src_lane = lane ^ ((pk >> 4) & 31);
          const int jsel = (pk >> 9) & 3;
#pragma unroll
          for (int j = 0; j < V; ++j) {
            double2 p = cur[j];
            if (LOG2V > 0) {
#pragma unroll
              for (int bit = 0; bit < LOG2V; ++bit) {
                const double2 cand = cur[j ^ (1 << bit)];
                const bool take = jsel == bit + 1;
                p.x = take ? cand.x : p.x; p.y = take ? cand.y : p.y;
              }
            }
            pv[j].x = __shfl_sync(FULL, p.x, src_lane);
            pv[j].y = __shfl_sync(FULL, p.y, src_lane);
          }
        } else {
#pragma unroll
          for (int j = 0; j < V; ++j) pv[j] = cur[j];
        }
        double ra, rb;
#pragma unroll
        for (int j = 0; j < V; ++j) {
          const unsigned i = li + L * j;
          const bool bw = i & wm, bc = i & cm;
          double2 v = cur[j];
          switch (kind) {
            case K_P:  if (bw) v = mul_phase(v, cs.x, cs.y); break;
            case K_CP: if (bw && bc) v = mul_phase(v, cs.x, cs.y); break;
            case K_RX: v = rot_rx(v, pv[j], cs.x, cs.y); break;
            case K_CX: if (bc) v = pv[j]; break;
            default: break;
          }
          if (j == 0) metric_set<UC>(v, tv[j], ra, rb); else metric_acc<UC>(v, tv[j], ra, rb);
          if (act) {
            if (reset) {
              if (fp) { float* f = fp + (long)(c0 + k) * gstride; f[L * j] = (float)v.x; f[S + L * j] = (float)v.y; }
              sp[L * j] = make_double2(rv[j], 0.0);
            } else if (kind != K_NONE) sp[L * j] = v;
          }
        }
        part[k * ROW + lane] = make_double2(ra, rb);
      }
      __syncwarp();
      {
        const int sidx = lane / P, prt = lane % P;
        const int k = sidx >> LOG2G, g2 = sidx & (G - 1);
        const double2* row = part + k * ROW + g2 * L + prt;
        double xa = 0.0, xb = 0.0;
#pragma unroll
        for (int jj = 0; jj < VAL; ++jj) {
          const double2 q = row[P * ((jj + g2) & (VAL - 1))];
          xa += q.x; xb += q.y;
        }
#pragma unroll
        for (int off = P / 2; off >= 1; off >>= 1) {
          xa += __shfl_xor_sync(FULL, xa, off);
          xb += __shfl_xor_sync(FULL, xb, off);
        }
        const int want = lane - c0 * G;
        const int from = (want & (SUMS - 1)) * P;
        const double va = __shfl_sync(FULL, xa, from), vb = __shfl_sync(FULL, xb, from);
        if (want >= 0 && want < SUMS) mysum = make_double2(va, vb);
      }
      __syncwarp();
    }
    phase3<UC, 32>(b, out, e, ctl, env, t, T, N, lane, mysum, ctl, stats);
    __syncwarp();                                 // s_cs / s_pk are rewritten by the next step
  }

  if (ctl) {
    reinterpret_cast<uint4*>(b.meta)[env] = e.meta;
    if (b.last) reinterpret_cast<double2*>(b.last)[env] = make_double2(e.last_m, e.last_c);
    if (b.ep_return) { b.ep_return[env] = e.ep_ret; b.ep_length[env] = e.ep_len; }
  }
  if (b.obs) {                                    // observation of the final states (env.py:85)
#pragma unroll 1
    for (int r = 0; r < R; ++r) {
      const int src = r * G + gi;
      if (src < n_valid) {
        float* o = b.obs + (env_base + src) * b.obs_stride + li;
#pragma unroll
        for (int j = 0; j < V; ++j) {
          const double2 v = stp[r * G * S + L * j];
          o[L * j] = (float)v.x; o[S + L * j] = (float)v.y;
        }
      }
    }
  }
  fence_async_smem();
  __syncwarp();
  if (lane == 0) {
    bulk_store(tstate, stile, (uint32_t)(n_valid * S * sizeof(double2)));
    bulk_store_wait_read();
  }
}

// ------------------------------------------------------------------------------------------------
// env.py:117-121 for the whole batch or a masked subset; one thread per amplitude.
__global__ void __launch_bounds__(256)
reset_kernel(const qcd_batch_t b, const uint8_t* __restrict__ mask, const int log2s) {
  const long S = 1L << log2s;
  const long gid = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const long env = gid >> log2s;
  if (env >= b.n_envs) return;
  if (mask && !mask[env]) return;
  const int i = (int)(gid & (S - 1));
  const bool uc = b.objective == QCD_OBJ_UC;
  const double2 v = reset_amp(uc, b.n_qubits, i);
  reinterpret_cast<double2*>(b.state)[env * S + i] = v;
  if (b.obs) {
    float* ob = b.obs + env * b.obs_stride;
    ob[i] = (float)v.x; ob[S + i] = 0.0f;
  }
  if (i == 0) {
    reinterpret_cast<uint4*>(b.meta)[env] = make_uint4(0, 0, 0, 0);
    if (b.last) reinterpret_cast<double2*>(b.last)[env] = make_double2(0.0, 0.0);
    if (b.ep_return) { b.ep_return[env] = 0.0; b.ep_length[env] = 0; }
    b.depth[env] = 0; b.operations[env] = 0; b.used_wires[env] = 0;       // info of env.py:86
    b.terminated[env] = 0; b.truncated[env] = 0; b.status[env] = 0;
    b.reward[env] = 0.0; b.metric[env] = 0.0; b.cost[env] = 0.0;
  }
}

// ------------------------------------------------------------------------------------------------
// Monitor episode fields (monitor.py:37-48) + termination histogram (algorithm.py:112-114) summed
// over the envs that finished in the last step.  Block reduce, one atomicAdd per field per CTA.
__global__ void __launch_bounds__(256)
stats_kernel(const qcd_batch_t b, double* __restrict__ stats) {
  __shared__ double s_part[8][QCD_STATS_LEN];
  double v[QCD_STATS_LEN];
#pragma unroll
  for (int k = 0; k < QCD_STATS_LEN; ++k) v[k] = 0.0;
  for (long env = (long)blockIdx.x * blockDim.x + threadIdx.x; env < b.n_envs;
       env += (long)gridDim.x * blockDim.x) {
    const int st = b.status[env];
    const bool term = b.terminated[env], trunc = b.truncated[env];
    if (st != QCD_ST_OK) { v[QCD_STAT_BAD_ACTIONS] += 1.0; continue; }
    v[QCD_STAT_STEPS] += 1.0;
    if (term || trunc) {
      v[QCD_STAT_EPISODES] += 1.0;
      if (b.episode_return) { v[QCD_STAT_SUM_R] += b.episode_return[env]; v[QCD_STAT_SUM_L] += (double)b.episode_length[env]; }
      v[QCD_STAT_SUM_D] += (double)b.depth[env];
      v[QCD_STAT_SUM_O] += (double)b.operations[env];
      v[QCD_STAT_SUM_Q] += (double)b.used_wires[env];
      v[QCD_STAT_SUM_M] += b.metric[env];
      v[QCD_STAT_SUM_C] += b.cost[env];
      if (trunc) v[QCD_STAT_N_DEPTH] += 1.0; else v[QCD_STAT_N_DONE] += 1.0;   // DEPTH overrides DONE (env.py:130-131)
    }
  }
#pragma unroll
  for (int k = 0; k < QCD_STATS_LEN; ++k) {
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], off);
  }
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int k = 0; k < QCD_STATS_LEN; ++k) s_part[threadIdx.x >> 5][k] = v[k];
  }
  __syncthreads();
  if (threadIdx.x < QCD_STATS_LEN) {
    double acc = 0.0;
    for (int wv = 0; wv < (int)(blockDim.x >> 5); ++wv) acc += s_part[wv][threadIdx.x];
    if (acc != 0.0) atomicAdd(stats + (blockIdx.x & (QCD_STATS_COPIES - 1)) * QCD_STATS_STRIDE + threadIdx.x, acc);
  }
}

// Sum of the QCD_STATS_COPIES replicas of the statistics vector in a fixed order -> out[QCD_STATS_STRIDE];
// optionally clears the replicas (start of the next reporting interval).  One CTA, one thread per replica
// element (all loads in flight at once: a serial loop over the replicas costs 64 dependent L2 round trips).
__global__ void __launch_bounds__(QCD_STATS_COPIES * QCD_STATS_STRIDE)
fold_stats_kernel(double* __restrict__ raw, double* __restrict__ out, const int zero_raw) {
  __shared__ double s_v[QCD_STATS_COPIES * QCD_STATS_STRIDE];
  pdl_wait();
  const int i = threadIdx.x;
  s_v[i] = raw[i];
  if (zero_raw) raw[i] = 0.0;
  __syncthreads();
  if (i < QCD_STATS_STRIDE) {
    double acc = 0.0;
#pragma unroll 8
    for (int c = 0; c < QCD_STATS_COPIES; ++c) acc += s_v[c * QCD_STATS_STRIDE + i];
    out[i] = i < QCD_STATS_LEN ? acc : 0.0;
  }
}

// ------------------------------------------------------------------------------------------------
// Rolling window of the most recent finished episodes (SB3's ep_info_buffer = deque(maxlen=100) that the
// reference's trainer averages: algorithm/algorithm.py:49,89-91,100), kept on the device.  The episodes that
// finished in the LAST step are appended in env order (the order DummyVecEnv reports them); only the last
// `window` of them are written, so the result does not depend on thread scheduling.  One CTA.
//   ring[w][QCD_EP_FIELDS] = {r, l, d, o, q, m, c, env}   count[0] = episodes pushed so far
__global__ void __launch_bounds__(1024)
push_episodes_kernel(const qcd_batch_t b, double* __restrict__ ring, long long* __restrict__ count, const int window) {
  __shared__ int s_warp[32];
  __shared__ int s_total;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long N = b.n_envs;
  auto finished = [&](const long env) {
    return env < N && b.status[env] == QCD_ST_OK && (b.terminated[env] || b.truncated[env]);
  };
  // pass 1: how many episodes finished in this step
  int mine = 0;
  for (long env = tid; env < N; env += 1024) mine += finished(env);
  mine = __reduce_add_sync(0xffffffffu, mine);
  if (tid == 0) s_total = 0;
  __syncthreads();
  if (lane == 0 && mine) atomicAdd(&s_total, mine);
  __syncthreads();
  const int total = s_total;
  if (total == 0) return;
  const long long base = count[0];
  const int skip = total > window ? total - window : 0;       // older than the window: never stored
  // pass 2: env-ordered rank of every finished episode (block scan per 1024 envs)
  int before = 0;
  for (long env0 = 0; env0 < N; env0 += 1024) {
    const long env = env0 + tid;
    const bool f = finished(env);
    const unsigned bal = __ballot_sync(0xffffffffu, f);
    if (lane == 0) s_warp[warp] = __popc(bal);
    __syncthreads();
    int wbase = 0, chunk = 0;
    for (int wv = 0; wv < 32; ++wv) { const int c = s_warp[wv]; if (wv < warp) wbase += c; chunk += c; }
    const int rank = before + wbase + __popc(bal & ((1u << lane) - 1u));
    if (f && rank >= skip) {
      double* row = ring + ((base + rank) % window) * QCD_EP_FIELDS;
      row[0] = b.episode_return ? b.episode_return[env] : 0.0;
      row[1] = b.episode_length ? (double)b.episode_length[env] : 0.0;
      row[2] = (double)b.depth[env]; row[3] = (double)b.operations[env]; row[4] = (double)b.used_wires[env];
      row[5] = b.metric[env]; row[6] = b.cost[env]; row[7] = (double)env;
    }
    before += chunk;
    __syncthreads();
  }
  if (tid == 0) count[0] = base + total;
}

#ifndef QCD_KERNELS_ONLY   // (a scratch translation unit may include this file for the kernels alone)
// ------------------------------------------------------------------------------------------------
// host-side dispatch
// The library keeps exactly two pieces of process-wide state, both idempotent and thread-safe: the
// QCD_PDL environment switch (read once) and, per kernel instantiation, a bit per device saying that the
// > 48 KiB dynamic shared memory opt-in has been made.
static bool pdl_enabled() {     // QCD_PDL=0 in the environment turns programmatic dependent launch off
  static const bool on = [] { const char* e = getenv("QCD_PDL"); return !(e && e[0] == '0'); }();
  return on;
}
struct SmemOptIn { std::atomic<uint64_t> done[4]; };   // 256 devices
template <typename Kern>
static int ensure_smem(Kern kern, const size_t smem, SmemOptIn& once) {
  if (smem == 0) return 0;
  int dev = 0;
  cudaError_t err = cudaGetDevice(&dev);
  if (err != cudaSuccess) return (int)err;
  const uint64_t bit = 1ull << (dev & 63);
  std::atomic<uint64_t>& word = once.done[(dev >> 6) & 3];
  if (word.load(std::memory_order_acquire) & bit) return 0;
  cudaFuncAttributes attr;                      // the 48 KiB default limit covers static + dynamic shared memory
  err = cudaFuncGetAttributes(&attr, kern);
  if (err != cudaSuccess) return (int)err;
  if (attr.sharedSizeBytes + smem > 48 * 1024) {
    err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return (int)err;
  }
  word.fetch_or(bit, std::memory_order_release);
  return 0;
}
constexpr int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
constexpr int ilog2(int v) { return v <= 1 ? 0 : 1 + ilog2(v >> 1); }

template <int LOG2S, bool UC, bool STAGED>
struct Launch {
  static constexpr int S = 1 << LOG2S;
  // step: <= one warp per env, streaming.  replay: a whole CTA per env once the state is >= 32 KiB.
  static constexpr int NT = STAGED ? (S >= 2048 ? 256 : 128) : (S <= QCD_PAIR_SMALL_MAX_S ? QCD_PAIR_SMALL_THREADS : 256);
  static constexpr int L = (STAGED && S >= 2048) ? NT : clampi(S / 4, 1, (!STAGED && S <= QCD_PAIR_NARROW_MAX_S) ? QCD_PAIR_NARROW_L : 32);
  static constexpr int LOG2L = ilog2(L);
  static constexpr int E = NT / L;
  static constexpr size_t SMEM = STAGED ? (size_t)E * S * sizeof(double2) : 0;
  static constexpr bool REG = !STAGED && LOG2S <= QCD_REG_STEP_MAX_LOG2S;   // register-resident step

  static int run(const qcd_batch_t& b, const void* actions, int act_f32, int T,
                 const qcd_tape_out_t& out, int metric_every_step, cudaStream_t stream, const double4 imm) {
    if constexpr (STAGED && S >= 512) {
      using M = RegMap<LOG2S, UC>;
      constexpr int MINB = 512 / M::NT;                       // 16 warps per SM at <= 128 registers
      auto kern = replay_reg_kernel<LOG2S, UC, MINB>;
      static SmemOptIn once;
      if (int rc = ensure_smem(kern, M::DYN_BYTES, once)) return rc;
      kern<<<(unsigned)b.n_envs, M::NT, M::DYN_BYTES, stream>>>(b, actions, act_f32, T, out, metric_every_step);
      return (int)cudaGetLastError();
    } else if constexpr (STAGED && S <= 256 && QCD_REPLAY_WARP) {
      if (b.flags & QCD_F_TARGET_PER_ENV) return run_pair(b, actions, act_f32, T, out, metric_every_step, stream, imm);
      constexpr int W = QCD_WARP_REPLAY_WARPS;
      constexpr int G = S < 32 ? 32 / S : 1;
      constexpr int TE0 = clampi(QCD_WARP_REPLAY_TILE_BYTES / (S * 16), 1, 32);
      constexpr int TE = TE0 < G ? G : TE0;
      using WC = WarpCfg<LOG2S, TE, true>;
      constexpr size_t smem = (size_t)W * WC::WARP_SLOTS * sizeof(double2);
      auto kern = replay_warp_kernel<LOG2S, UC, W, TE, QCD_WARP_REPLAY_MINB>;
      static SmemOptIn once;
      if (int rc = ensure_smem(kern, smem, once)) return rc;
      const long tiles = ((long)b.n_envs + TE - 1) / TE;
      const unsigned grid = (unsigned)((tiles + W - 1) / W);
      kern<<<grid, W * 32, smem, stream>>>(b, actions, act_f32, T, out, metric_every_step);
      return (int)cudaGetLastError();
    } else if constexpr (REG && QCD_STEP_IMPL == 2) {
      if (b.flags & QCD_F_TARGET_PER_ENV) return run_pair(b, actions, act_f32, T, out, metric_every_step, stream, imm);
      constexpr int W = QCD_WARP_STEP_WARPS;
      constexpr int G = S < 32 ? 32 / S : 1;
      constexpr int TE0 = QCD_WARP_STEP_TILE > 0 ? QCD_WARP_STEP_TILE : clampi(8192 / (S * 16), 8, 32);
      constexpr int TE = TE0 < G ? G : TE0;
      constexpr bool STG = QCD_WARP_STEP_STAGE && (TE * S * 16 <= 16384);
      using WC = WarpCfg<LOG2S, TE, STG>;
      constexpr size_t smem = (size_t)W * WC::WARP_SLOTS * sizeof(double2);
      auto kern = step_warp_kernel<LOG2S, UC, W, TE, STG, QCD_WARP_STEP_MINB>;
      static SmemOptIn once;
      if (int rc = ensure_smem(kern, smem, once)) return rc;
      const long tiles = ((long)b.n_envs + TE - 1) / TE;
      const unsigned grid = (unsigned)((tiles + W - 1) / W);
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(grid); cfg.blockDim = dim3(W * 32); cfg.dynamicSmemBytes = smem; cfg.stream = stream;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      attr[0].val.programmaticStreamSerializationAllowed = 1;
      cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 1 : 0;
      return (int)cudaLaunchKernelEx(&cfg, kern, b, actions, act_f32, imm);
    } else {
      return run_pair(b, actions, act_f32, T, out, metric_every_step, stream, imm);
    }
  }

  static int run_pair(const qcd_batch_t& b, const void* actions, int act_f32, int T,
                      const qcd_tape_out_t& out, int metric_every_step, cudaStream_t stream, const double4 imm) {
    auto kern = qcd_kernel<LOG2S, UC, STAGED, LOG2L, NT>;
    static SmemOptIn once;
    if (int rc = ensure_smem(kern, SMEM, once)) return rc;
    const unsigned grid = (unsigned)((b.n_envs + E - 1) / E);
    kern<<<grid, NT, SMEM, stream>>>(b, actions, act_f32, T, out, metric_every_step, imm);
    return (int)cudaGetLastError();
  }
};

template <bool STAGED>
int dispatch(const qcd_batch_t& b, const void* actions, int act_f32, int T, const qcd_tape_out_t& out,
             int mes, cudaStream_t s, const double4 imm = double4{0.0, 0.0, 0.0, 0.0}) {
  const bool uc = b.objective == QCD_OBJ_UC;
  const int log2s = uc ? 2 * b.n_qubits : b.n_qubits;
#define QCD_CASE(LS)                                                                   \
  case LS:                                                                             \
    return uc ? Launch<LS, true, STAGED>::run(b, actions, act_f32, T, out, mes, s, imm) \
              : Launch<LS, false, STAGED>::run(b, actions, act_f32, T, out, mes, s, imm);
#define QCD_CASE_SP(LS) \
  case LS:              \
    return uc ? QCD_E_RANGE : Launch<LS, false, STAGED>::run(b, actions, act_f32, T, out, mes, s, imm);
  switch (log2s) {
    QCD_CASE_SP(1) QCD_CASE(2) QCD_CASE_SP(3) QCD_CASE(4) QCD_CASE_SP(5) QCD_CASE(6)
    QCD_CASE_SP(7) QCD_CASE(8) QCD_CASE_SP(9) QCD_CASE(10) QCD_CASE_SP(11) QCD_CASE(12)
    default: return QCD_E_RANGE;
  }
#undef QCD_CASE
#undef QCD_CASE_SP
}

int validate(const qcd_batch_t* b, bool need_actions, const void* actions, int action_dtype) {
  if (!b) return QCD_E_NULL;
  if (b->objective != QCD_OBJ_SP && b->objective != QCD_OBJ_UC) return QCD_E_OBJ;
  if (action_dtype != QCD_ACT_F64 && action_dtype != QCD_ACT_F32) return QCD_E_OBJ;
  const int maxq = b->objective == QCD_OBJ_UC ? QCD_MAX_QUBITS_UC : QCD_MAX_QUBITS_SP;
  if (b->n_envs < 1 || b->n_qubits < 1 || b->n_qubits > maxq) return QCD_E_RANGE;
  if (b->max_depth < 1 || b->max_depth > QCD_MAX_DEPTH) return QCD_E_RANGE;
  const int64_t S = b->objective == QCD_OBJ_UC ? (1LL << (2 * b->n_qubits)) : (1LL << b->n_qubits);
  if ((b->obs || b->final_obs) && b->obs_stride < 2 * S) return QCD_E_STRIDE;
  if (!b->state || !b->target || !b->meta || !b->reward || !b->metric || !b->cost || !b->depth ||
      !b->operations || !b->used_wires || !b->terminated || !b->truncated || !b->status)
    return QCD_E_NULL;
  if (!(b->flags & QCD_F_SPARSE) && !b->last) return QCD_E_NULL;
  if ((b->ep_return == nullptr) != (b->ep_length == nullptr)) return QCD_E_NULL;
  if ((b->episode_return == nullptr) != (b->episode_length == nullptr)) return QCD_E_NULL;
  if (b->episode_return && !b->ep_return) return QCD_E_NULL;
  if (need_actions && !actions) return QCD_E_NULL;
  return 0;
}

}  // namespace qcd

// =================================================================================================
extern "C" {

int qcd_abi_version(void) { return QCD_ABI_VERSION; }

int qcd_reset(const qcd_batch_t* b, const uint8_t* mask, void* stream) {
  if (b && b->n_envs == 0) return 0;          // empty batch: nothing to do
  int rc = qcd::validate(b, false, nullptr, QCD_ACT_F64);
  if (rc) return rc;
  const int log2s = b->objective == QCD_OBJ_UC ? 2 * b->n_qubits : b->n_qubits;
  const long total = (long)b->n_envs << log2s;
  const unsigned grid = (unsigned)((total + 255) / 256);
  qcd::reset_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(*b, mask, log2s);
  return (int)cudaGetLastError();
}

int qcd_step(const qcd_batch_t* b, const void* actions, int action_dtype, void* stream) {
  if (b && b->n_envs == 0) return 0;          // empty batch: nothing to do
  int rc = qcd::validate(b, true, actions, action_dtype);
  if (rc) return rc;
  const qcd_tape_out_t none = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  return qcd::dispatch<false>(*b, actions, action_dtype == QCD_ACT_F32, 1, none, 1, (cudaStream_t)stream);
}

int qcd_replay(const qcd_batch_t* b, const void* actions, int action_dtype, int32_t n_steps,
               const qcd_tape_out_t* out, void* stream) {
  if (b && b->n_envs == 0) return 0;          // empty batch: nothing to do
  int rc = qcd::validate(b, true, actions, action_dtype);
  if (rc) return rc;
  if (n_steps < 1) return QCD_E_RANGE;
  const qcd_tape_out_t none = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  const qcd_tape_out_t o = out ? *out : none;
  // info['metric'] is defined on every step (env.py:133); it is only *needed* on every step when
  // someone records it or when the reward is dense.
  const int every = (o.metric != nullptr) || !(b->flags & QCD_F_SPARSE);
  return qcd::dispatch<true>(*b, actions, action_dtype == QCD_ACT_F32, n_steps, o, every, (cudaStream_t)stream);
}

int qcd_reduce_stats(const qcd_batch_t* b, double* stats, void* stream) {
  if (b && b->n_envs == 0) return 0;          // empty batch: nothing to do
  int rc = qcd::validate(b, false, nullptr, QCD_ACT_F64);
  if (rc) return rc;
  if (!stats) return QCD_E_NULL;
  unsigned grid = (unsigned)((b->n_envs + 255) / 256);
  if (grid > 148 * 8) grid = 148 * 8;
  qcd::stats_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(*b, stats);
  return (int)cudaGetLastError();
}

int qcd_fold_stats(double* stats_raw, double* out, int zero_raw, void* stream) {
  if (!stats_raw || !out) return QCD_E_NULL;
  // launched like qcd_step (programmatic stream serialization): its launch latency hides behind the tail of the
  // step before it; griddepcontrol.wait at its top orders it after that step's statistics atomics
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(1); cfg.blockDim = dim3(QCD_STATS_COPIES * QCD_STATS_STRIDE); cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = qcd::pdl_enabled() ? 1 : 0;
  return (int)cudaLaunchKernelEx(&cfg, qcd::fold_stats_kernel, stats_raw, out, zero_raw);
}

int qcd_push_episodes(const qcd_batch_t* b, double* ring, int64_t* count, int32_t window, void* stream) {
  if (b && b->n_envs == 0) return 0;
  int rc = qcd::validate(b, false, nullptr, QCD_ACT_F64);
  if (rc) return rc;
  if (!ring || !count) return QCD_E_NULL;
  if (window < 1) return QCD_E_RANGE;
  qcd::push_episodes_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(*b, ring, reinterpret_cast<long long*>(count), window);
  return (int)cudaGetLastError();
}

int qcd_step_sync(const qcd_batch_t* b, const void* actions, int action_dtype, void* stream) {
  int rc = qcd_step(b, actions, action_dtype, stream);
  if (rc) return rc;
  return (int)cudaStreamSynchronize((cudaStream_t)stream);
}

int qcd_step1_sync(const qcd_batch_t* b, const double* action, void* stream) {
  int rc = qcd::validate(b, false, nullptr, QCD_ACT_F64);
  if (rc) return rc;
  if (!action) return QCD_E_NULL;
  if (b->n_envs != 1) return QCD_E_RANGE;
  const qcd_tape_out_t none = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  const double4 imm = {action[0], action[1], action[2], action[3]};      // travels as a kernel argument
  rc = qcd::dispatch<false>(*b, nullptr, qcd::ACT_IMMEDIATE, 1, none, 1, (cudaStream_t)stream, imm);
  if (rc) return rc;
  return (int)cudaStreamSynchronize((cudaStream_t)stream);
}

int qcd_step_host(const qcd_batch_t* b, const void* actions_host, void* actions_dev, int action_dtype,
                  const qcd_host_out_t* out, void* stream) {
  if (b && b->n_envs == 0) return 0;          // empty batch: nothing to do
  int rc = qcd::validate(b, true, actions_dev, action_dtype);
  if (rc) return rc;
  if (!actions_host || !out) return QCD_E_NULL;
  cudaStream_t s = (cudaStream_t)stream;
  const size_t n = (size_t)b->n_envs;
  const size_t abytes = n * 4 * (action_dtype == QCD_ACT_F32 ? 4 : 8);
  const int64_t S = b->objective == QCD_OBJ_UC ? (1LL << (2 * b->n_qubits)) : (1LL << b->n_qubits);
  if (out->obs) {
    if (!b->obs) return QCD_E_NULL;
    if (out->obs_stride < 2 * S) return QCD_E_STRIDE;
  }
  cudaError_t err = cudaMemcpyAsync(actions_dev, actions_host, abytes, cudaMemcpyHostToDevice, s);
  if (err != cudaSuccess) return (int)err;
  rc = qcd_step(b, actions_dev, action_dtype, stream);
  if (rc) return rc;
  // the bulk of the bytes first: the state half of every observation row.  With compact rows on both
  // sides (stride 2*S: obs_mode 'state') this is ONE contiguous transfer; otherwise a pitched one.
  if (out->obs) {
    if (out->obs_stride == 2 * S && b->obs_stride == 2 * S)
      err = cudaMemcpyAsync(out->obs, b->obs, n * (size_t)(2 * S) * sizeof(float), cudaMemcpyDeviceToHost, s);
    else
      err = cudaMemcpy2DAsync(out->obs, (size_t)out->obs_stride * sizeof(float), b->obs,
                              (size_t)b->obs_stride * sizeof(float), (size_t)(2 * S) * sizeof(float), n,
                              cudaMemcpyDeviceToHost, s);
    if (err != cudaSuccess) return (int)err;
  }
  const struct { void* dst; const void* src; size_t bytes; } copies[] = {
      {out->reward, b->reward, n * sizeof(double)},        {out->metric, b->metric, n * sizeof(double)},
      {out->cost, b->cost, n * sizeof(double)},            {out->depth, b->depth, n * sizeof(int32_t)},
      {out->operations, b->operations, n * sizeof(int32_t)}, {out->used_wires, b->used_wires, n * sizeof(int32_t)},
      {out->terminated, b->terminated, n},                 {out->truncated, b->truncated, n},
      {out->status, b->status, n}};
  for (const auto& c : copies) {
    if (!c.dst) continue;
    err = cudaMemcpyAsync(c.dst, c.src, c.bytes, cudaMemcpyDeviceToHost, s);
    if (err != cudaSuccess) return (int)err;
  }
  return 0;
}

}  // extern "C"
#endif  // QCD_KERNELS_ONLY

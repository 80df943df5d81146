// qcd_kernels.cu — batched CircuitDesigner hot path for B200 (sm_100a).
//
// One kernel template, two instantiation families:
//   STAGED = false : the STEP kernel.  One launch = one env.step() of every env.  State is streamed
//                    HBM -> registers -> HBM exactly once; the kernel is HBM-bound (32*S + 8*S + ~96
//                    bytes per env-step, S = 2^eta (SP) or 4^eta (UC) complex128 amplitudes).
//   STAGED = true  : the REPLAY kernel.  Each env's state is loaded into shared memory once, T actions
//                    are applied in place, and it is written back once (FP64-pipe / smem bound).
//
// Reference semantics reproduced (file:line in /root/reference):
//   qcd_gym/env.py:90-100   action decode <o,q,c,Phi>            -> decode_action()
//   qcd_gym/env.py:83-84    gate applied to Statevector/Operator -> apply_pair() (incremental: the
//                            reference re-simulates the whole circuit, same ops in the same order)
//   qcd_gym/env.py:86,129   depth / operations / used wires, truncation -> decode_action()
//   qcd_gym/env.py:106-114  fidelity | arctan-Frobenius metric, depth cost -> finalize_step()
//   qcd_gym/env.py:134-135  sparse / punish shaping                -> finalize_step()
//   qcd_gym/env.py:10,85    float32 observation [Re | Im]          -> store phase of run_phase2()
//   qcd_gym/env.py:117-121  reset                                   -> fused auto-reset + reset_kernel
//   qcd_gym/wrappers/monitor.py:37-48 episode return/length        -> finalize_step(), stats_kernel
//
// Mapping.  A gate on qubit w only couples amplitude pairs (i0, i1 = i0 | 1<<w).  Every gate of the
// reference's set {P, RX, CP, CX} is a 2x2 butterfly on such a pair (CP/CX predicated on the control
// bit of i0), so a group of L = 2^LOG2L threads owns one env and each thread owns PPT = S/2/L pairs;
// pair k maps to i0 by inserting a zero bit at position w.  No thread ever needs another thread's
// amplitudes, gates are heterogeneous per env for free (only the address pattern and two doubles
// differ), and UC is the same butterfly on the row bits (w + eta) of the row-major unitary.
// The per-env scalar work (floor-decode, sincos, depth stack, cost, reward) is done by ONE thread per
// env with consecutive threads on consecutive envs (fully populated warps, coalesced scalar I/O) and
// handed to the group through shared memory.

#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include "qcd_b200.h"

namespace qcd {

constexpr int K_NONE = 0, K_P = 1, K_RX = 2, K_CP = 3, K_CX = 4;

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

// env.py:90-100 (decode) + QuantumCircuit.depth/count_ops/idle_wires bookkeeping (env.py:74,77,86)
// + truncation (env.py:129) + cost (env.py:112).  Integer results are bit-exact by construction:
// floor() is taken in float64 on exactly-widened inputs and the range test is done before the cast.
__device__ __forceinline__ void decode_action(const qcd_batch_t& b, double a0, double a1, double a2,
                                              double a3, EnvCtl& e, GateRec& rec) {
  const double fg = floor(a0), fw = floor(a1), fc = floor(a2);
  const double eta = (double)b.n_qubits;
  int status = QCD_ST_OK;
  if (!(fw >= 0.0 && fw < eta && fc >= 0.0 && fc < eta)) status = QCD_ST_BAD_WIRE;     // env.py:94
  else if (!(fg == 0.0 || fg == 1.0 || fg == 2.0)) status = QCD_ST_BAD_GATE;           // env.py:100
  int kind = K_NONE, w = 0, c = 0;
  bool term = false;
  if (status == QCD_ST_OK) {
    w = (int)fw; c = (int)fc;
    const int g = (int)fg;
    if (g == 2) term = true;                                   // env.py:99 (checked after 95-98,
    else if (w == c) kind = (g == 0) ? K_P : K_RX;             //  but g==2 matches none of those)
    else kind = (g == 0) ? K_CP : K_CX;
  }
  double cr = 1.0, sr = 0.0;
  if (kind == K_P || kind == K_CP || kind == K_RX) {
    const double arg = (kind == K_RX) ? a3 / 2 : a3;           // RXGate: cos/sin(theta/2)
    sincos(arg, &sr, &cr);
  }
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
  e.cost_raw = fmax(0.0, (double)depth - b.depth_free) / b.cost_denom;                 // env.py:112
  rec.cr = cr; rec.sr = sr; rec.kind = kind;
  const int shift = (b.objective == QCD_OBJ_UC) ? b.n_qubits : 0;   // UC: gate acts on the row index
  rec.wbit = w + shift; rec.cbit = c + shift;
  rec.flags = ((e.term || e.trunc) && (b.flags & QCD_F_AUTORESET)) ? RF_RESET : 0;
}

// env.py:106-114 + 133-135 + Monitor accumulators.  `ra`,`rb` are the reduced metric sums.
struct StepOut { double reward, metric, cost; };

template <bool UC>
__device__ __forceinline__ StepOut finalize_step(const qcd_batch_t& b, EnvCtl& e, double ra, double rb,
                                                 bool have_metric) {
  double metric_raw;
  if (UC) {
    const double nrm = sqrt(ra);                               // ||U - V||_F
    metric_raw = 1.0 - 2.0 * atan(nrm) / CUDART_PI;            // env.py:111
  } else {
    const double h = hypot(ra, rb);                            // abs(vdot)
    metric_raw = h * h;                                        // env.py:109
  }
  if (!have_metric) metric_raw = CUDART_NAN;
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

__device__ __forceinline__ double2 reset_amp(bool uc, int eta, int i) {
  const bool one = uc ? ((i >> eta) == (i & ((1 << eta) - 1))) : (i == 0);
  return make_double2(one ? 1.0 : 0.0, 0.0);
}

// The 2x2 butterfly on one amplitude pair; (a0,a1) at (i0, i0|1<<wbit).
__device__ __forceinline__ void apply_pair(const GateRec& g, int i0, double2& a0, double2& a1) {
  switch (g.kind) {
    case K_P: {                       // diag(1, e^{i th}) on the w-bit
      const double re = a1.x * g.cr - a1.y * g.sr, im = a1.x * g.sr + a1.y * g.cr;
      a1.x = re; a1.y = im;
    } break;
    case K_CP: {                      // diag(1,1,1,e^{i th}): both bits set
      if ((i0 >> g.cbit) & 1) {
        const double re = a1.x * g.cr - a1.y * g.sr, im = a1.x * g.sr + a1.y * g.cr;
        a1.x = re; a1.y = im;
      }
    } break;
    case K_RX: {                      // [[c, -is], [-is, c]]
      const double r0 = g.cr * a0.x + g.sr * a1.y, m0 = g.cr * a0.y - g.sr * a1.x;
      const double r1 = g.sr * a0.y + g.cr * a1.x, m1 = g.cr * a1.y - g.sr * a0.x;
      a0.x = r0; a0.y = m0; a1.x = r1; a1.y = m1;
    } break;
    case K_CX: {                      // swap the pair where the control bit is set
      if ((i0 >> g.cbit) & 1) { const double2 t = a0; a0 = a1; a1 = t; }
    } break;
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
           const qcd_tape_out_t out, const int metric_every_step) {
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
  const bool autoreset = b.flags & QCD_F_AUTORESET;

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

  if (STAGED) {                                   // stage the CTA's contiguous block of states
    const long base = env0 * S;
    const long lim = (N - env0 < E ? N - env0 : E) * (long)S;
    for (long i = tid; i < lim; i += NT) s_tile[i] = gstate[base + i];
  }

  for (int t = 0; t < T; ++t) {
    // ================= phase 1: per-env scalar decode
    if (ctl) {
      double a0, a1, a2, a3;
      load_action(actions, act_f32, (long)t * N + cenv, a0, a1, a2, a3);
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
      float* fob = (b.final_obs && last && reset && !STAGED) ? b.final_obs + genv * b.obs_stride : nullptr;
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
          if (wr_state) { st[i0[u]] = x0[u]; st[i1] = x1[u]; }
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

    // ================= phase 3: per-env finalize + scalar outputs
    int st_fin = 0, st_ok = 0, st_bad = 0, st_trunc = 0, st_l = 0, st_d = 0, st_o = 0, st_q = 0;
    double st_r = 0.0, st_m = 0.0, st_c = 0.0;
    if (ctl) {
      const double2 r = s_red[tid];
      const bool have_metric = s_rec[tid].flags & RF_METRIC;
      const StepOut o = finalize_step<UC>(b, e, r.x, r.y, have_metric);
      const bool done = (e.status == QCD_ST_OK) && (e.term || e.trunc);
      if (out.reward) out.reward[(long)t * N + cenv] = o.reward;
      if (out.metric) out.metric[(long)t * N + cenv] = o.metric;
      if (out.cost) out.cost[(long)t * N + cenv] = o.cost;
      if (out.depth) out.depth[(long)t * N + cenv] = e.depth;
      if (out.terminated) out.terminated[(long)t * N + cenv] = e.term;
      if (out.truncated) out.truncated[(long)t * N + cenv] = e.trunc;
      if (out.status) out.status[(long)t * N + cenv] = (int8_t)e.status;
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
    if (b.stats != nullptr && (tid >> 5) <= ((E - 1) >> 5)) {
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
          atomicAdd(b.stats + QCD_STAT_EPISODES, (double)n_fin);
          if (b.ep_return) { atomicAdd(b.stats + QCD_STAT_SUM_R, st_r); atomicAdd(b.stats + QCD_STAT_SUM_L, (double)s_l); }
          atomicAdd(b.stats + QCD_STAT_SUM_D, (double)s_d);
          atomicAdd(b.stats + QCD_STAT_SUM_O, (double)s_o);
          atomicAdd(b.stats + QCD_STAT_SUM_Q, (double)s_q);
          atomicAdd(b.stats + QCD_STAT_SUM_M, st_m);
          atomicAdd(b.stats + QCD_STAT_SUM_C, st_c);
          if (n_fin - n_trunc) atomicAdd(b.stats + QCD_STAT_N_DONE, (double)(n_fin - n_trunc));
          if (n_trunc) atomicAdd(b.stats + QCD_STAT_N_DEPTH, (double)n_trunc);
        }
      }
      if ((tid & 31) == 0) {
        if (n_ok) atomicAdd(b.stats + QCD_STAT_STEPS, (double)n_ok);
        if (n_bad) atomicAdd(b.stats + QCD_STAT_BAD_ACTIONS, (double)n_bad);
      }
    }
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
    if (acc != 0.0) atomicAdd(stats + threadIdx.x, acc);
  }
}

// ------------------------------------------------------------------------------------------------
// host-side dispatch
constexpr int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
constexpr int ilog2(int v) { return v <= 1 ? 0 : 1 + ilog2(v >> 1); }

template <int LOG2S, bool UC, bool STAGED>
struct Launch {
  static constexpr int S = 1 << LOG2S;
  // step: <= one warp per env, streaming.  replay: a whole CTA per env once the state is >= 32 KiB.
  static constexpr int NT = STAGED ? (S >= 2048 ? 256 : 128) : 256;
  static constexpr int L = (STAGED && S >= 2048) ? NT : clampi(S / 4, 1, 32);
  static constexpr int LOG2L = ilog2(L);
  static constexpr int E = NT / L;
  static constexpr size_t SMEM = STAGED ? (size_t)E * S * sizeof(double2) : 0;

  static int run(const qcd_batch_t& b, const void* actions, int act_f32, int T,
                 const qcd_tape_out_t& out, int metric_every_step, cudaStream_t stream) {
    auto kern = qcd_kernel<LOG2S, UC, STAGED, LOG2L, NT>;
    if (SMEM > 48 * 1024) {
      cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM);
      if (err != cudaSuccess) return (int)err;
    }
    const unsigned grid = (unsigned)((b.n_envs + E - 1) / E);
    kern<<<grid, NT, SMEM, stream>>>(b, actions, act_f32, T, out, metric_every_step);
    return (int)cudaGetLastError();
  }
};

template <bool STAGED>
int dispatch(const qcd_batch_t& b, const void* actions, int act_f32, int T, const qcd_tape_out_t& out,
             int mes, cudaStream_t s) {
  const bool uc = b.objective == QCD_OBJ_UC;
  const int log2s = uc ? 2 * b.n_qubits : b.n_qubits;
#define QCD_CASE(LS)                                                                   \
  case LS:                                                                             \
    return uc ? Launch<LS, true, STAGED>::run(b, actions, act_f32, T, out, mes, s)     \
              : Launch<LS, false, STAGED>::run(b, actions, act_f32, T, out, mes, s);
#define QCD_CASE_SP(LS) \
  case LS:              \
    return uc ? QCD_E_RANGE : Launch<LS, false, STAGED>::run(b, actions, act_f32, T, out, mes, s);
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
  int rc = qcd::validate(b, false, nullptr, QCD_ACT_F64);
  if (rc) return rc;
  const int log2s = b->objective == QCD_OBJ_UC ? 2 * b->n_qubits : b->n_qubits;
  const long total = (long)b->n_envs << log2s;
  const unsigned grid = (unsigned)((total + 255) / 256);
  qcd::reset_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(*b, mask, log2s);
  return (int)cudaGetLastError();
}

int qcd_step(const qcd_batch_t* b, const void* actions, int action_dtype, void* stream) {
  int rc = qcd::validate(b, true, actions, action_dtype);
  if (rc) return rc;
  const qcd_tape_out_t none = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  return qcd::dispatch<false>(*b, actions, action_dtype == QCD_ACT_F32, 1, none, 1, (cudaStream_t)stream);
}

int qcd_replay(const qcd_batch_t* b, const void* actions, int action_dtype, int32_t n_steps,
               const qcd_tape_out_t* out, void* stream) {
  int rc = qcd::validate(b, true, actions, action_dtype);
  if (rc) return rc;
  if (n_steps < 1) return QCD_E_RANGE;
  const qcd_tape_out_t none = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  const qcd_tape_out_t o = out ? *out : none;
  // info['metric'] is defined on every step (env.py:133); it is only *needed* on every step when
  // someone records it or when the reward is dense.
  const int every = (o.metric != nullptr) || !(b->flags & QCD_F_SPARSE);
  return qcd::dispatch<true>(*b, actions, action_dtype == QCD_ACT_F32, n_steps, o, every, (cudaStream_t)stream);
}

int qcd_reduce_stats(const qcd_batch_t* b, double* stats, void* stream) {
  int rc = qcd::validate(b, false, nullptr, QCD_ACT_F64);
  if (rc) return rc;
  if (!stats) return QCD_E_NULL;
  unsigned grid = (unsigned)((b->n_envs + 255) / 256);
  if (grid > 148 * 8) grid = 148 * 8;
  qcd::stats_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(*b, stats);
  return (int)cudaGetLastError();
}

int qcd_step_host(const qcd_batch_t* b, const void* actions_host, void* actions_dev, int action_dtype,
                  double* reward_host, uint8_t* terminated_host, uint8_t* truncated_host,
                  float* obs_host, int64_t obs_host_stride, void* stream) {
  int rc = qcd::validate(b, true, actions_dev, action_dtype);
  if (rc) return rc;
  if (!actions_host || !reward_host || !terminated_host || !truncated_host) return QCD_E_NULL;
  cudaStream_t s = (cudaStream_t)stream;
  const size_t abytes = (size_t)b->n_envs * 4 * (action_dtype == QCD_ACT_F32 ? 4 : 8);
  cudaError_t err = cudaMemcpyAsync(actions_dev, actions_host, abytes, cudaMemcpyHostToDevice, s);
  if (err != cudaSuccess) return (int)err;
  rc = qcd_step(b, actions_dev, action_dtype, stream);
  if (rc) return rc;
  const size_t n = (size_t)b->n_envs;
  err = cudaMemcpyAsync(reward_host, b->reward, n * sizeof(double), cudaMemcpyDeviceToHost, s);
  if (err != cudaSuccess) return (int)err;
  err = cudaMemcpyAsync(terminated_host, b->terminated, n, cudaMemcpyDeviceToHost, s);
  if (err != cudaSuccess) return (int)err;
  err = cudaMemcpyAsync(truncated_host, b->truncated, n, cudaMemcpyDeviceToHost, s);
  if (err != cudaSuccess) return (int)err;
  if (obs_host) {
    if (!b->obs) return QCD_E_NULL;
    const int64_t S = b->objective == QCD_OBJ_UC ? (1LL << (2 * b->n_qubits)) : (1LL << b->n_qubits);
    if (obs_host_stride < 2 * S) return QCD_E_STRIDE;
    err = cudaMemcpy2DAsync(obs_host, (size_t)obs_host_stride * sizeof(float), b->obs,
                            (size_t)b->obs_stride * sizeof(float), (size_t)(2 * S) * sizeof(float), n,
                            cudaMemcpyDeviceToHost, s);
    if (err != cudaSuccess) return (int)err;
  }
  return 0;
}

}  // extern "C"

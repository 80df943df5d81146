/* qcd_oracle_batch.c — plain-C restatement of the BATCHED step semantics (TEST INFRASTRUCTURE).
 *
 * Not product code: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load it.
 * It restates, per env, /root/reference/qcd_gym/env.py:90-136 with the incremental (one gate per
 * step) evolution that is arithmetically the same operation sequence as the reference's
 * from-scratch re-simulation, plus the VectorEnv same-step auto-reset and the Monitor
 * accumulators (qcd_gym/wrappers/monitor.py:37-38).  Its own correctness is pinned in
 * tests/test_oracle_batch.py against oracle/qcd_oracle.py (the einsum restatement of qiskit 1.0.2),
 * which in turn is pinned on the reference's four KATs.  Parity status: see qcd_oracle.py header.
 *
 * Scalar, single-threaded, -ffp-contract=off: every product/sum is rounded separately, like numpy.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

#define F_PUNISH 1u
#define F_SPARSE 2u
#define F_AUTORESET 4u
#define F_TARGET_PER_ENV 8u

typedef struct {
  int32_t n_envs, n_qubits, objective, max_depth;
  uint32_t flags; int32_t reserved0;
  double depth_free, cost_denom;
  int64_t obs_stride;
  double* state; const double* target; uint8_t* meta; double* last;
  double* ep_return; int32_t* ep_length;
  float* obs; float* final_obs;
  double* reward; double* metric; double* cost;
  int32_t* depth; int32_t* operations; int32_t* used_wires;
  uint8_t* terminated; uint8_t* truncated; int8_t* status;
  double* episode_return; int32_t* episode_length;
  double* stats;   /* [12] running episode statistics (monitor.py:37-48, algorithm.py:112-114) or NULL */
} batch_t;

static void write_obs(float* row, const double* st, long S) {   /* env.py:10 flat(): [Re | Im] float32 */
  for (long i = 0; i < S; ++i) { row[i] = (float)st[2 * i]; row[S + i] = (float)st[2 * i + 1]; }
}

static void reset_state(double* st, long S, int uc, int eta) {   /* e0 or identity */
  memset(st, 0, sizeof(double) * 2 * (size_t)S);
  if (!uc) { st[0] = 1.0; return; }
  long D = 1L << eta;
  for (long r = 0; r < D; ++r) st[2 * (r * D + r)] = 1.0;
}

int oracle_reset(const batch_t* b, const uint8_t* mask) {
  const int uc = b->objective == 1;
  const long S = 1L << (uc ? 2 * b->n_qubits : b->n_qubits);
  for (long e = 0; e < b->n_envs; ++e) {
    if (mask && !mask[e]) continue;
    double* st = b->state + 2 * S * e;
    reset_state(st, S, uc, b->n_qubits);
    memset(b->meta + 16 * e, 0, 16);
    if (b->last) { b->last[2 * e] = 0; b->last[2 * e + 1] = 0; }
    if (b->ep_return) { b->ep_return[e] = 0; b->ep_length[e] = 0; }
    if (b->obs) write_obs(b->obs + e * b->obs_stride, st, S);
    b->depth[e] = 0; b->operations[e] = 0; b->used_wires[e] = 0;
    b->terminated[e] = 0; b->truncated[e] = 0; b->status[e] = 0;
    b->reward[e] = 0; b->metric[e] = 0; b->cost[e] = 0;
  }
  return 0;
}

static void cmul(double* a, double cr, double ci) {  /* a *= (cr + i ci), numpy complex product */
  const double re = a[0] * cr - a[1] * ci, im = a[0] * ci + a[1] * cr;
  a[0] = re; a[1] = im;
}

/* one env.step() for every env; actions float64 [N,4] */
int oracle_step(const batch_t* b, const double* actions) {
  const int uc = b->objective == 1, eta = b->n_qubits;
  const long S = 1L << (uc ? 2 * eta : eta);
  const int shift = uc ? eta : 0;
  for (long e = 0; e < b->n_envs; ++e) {
    const double* a = actions + 4 * e;
    double* st = b->state + 2 * S * e;
    const double* tg = b->target + ((b->flags & F_TARGET_PER_ENV) ? 2 * S * e : 0);
    uint8_t* meta = b->meta + 16 * e;
    uint16_t ops, used;
    memcpy(&ops, meta + 12, 2); memcpy(&used, meta + 14, 2);

    /* ---- env.py:90-100 */
    const double fg = floor(a[0]), fw = floor(a[1]), fc = floor(a[2]), theta = a[3];
    int status = 0, term = 0, kind = 0, w = 0, c = 0;   /* kind: 1 P, 2 RX, 3 CP, 4 CX */
    if (!(fw >= 0 && fw < eta && fc >= 0 && fc < eta)) status = 1;
    else if (!(fg == 0 || fg == 1 || fg == 2)) status = 2;
    else {
      w = (int)fw; c = (int)fc;
      const int g = (int)fg;
      if (w == c && g == 0) kind = 1;
      else if (w == c && g == 1) kind = 2;
      else if (g == 0) kind = 3;
      else if (g == 1) kind = 4;
      else term = 1;
    }
    /* ---- env.py:127 append + qiskit evolution of this one gate */
    if (kind) {
      const long bw = 1L << (w + shift), bc = 1L << (c + shift);
      if (kind == 1 || kind == 3) {
        const double cr = cos(theta), ci = sin(theta);          /* cmath.exp(1j*theta) */
        for (long i = 0; i < S; ++i)
          if ((i & bw) && (kind == 1 || (i & bc))) cmul(st + 2 * i, cr, ci);
      } else if (kind == 2) {
        const double cs = cos(theta / 2), sn = sin(theta / 2);
        for (long i = 0; i < S; ++i) {
          if (i & bw) continue;
          double* a0 = st + 2 * i; double* a1 = st + 2 * (i | bw);
          /* [[cs, -i sn], [-i sn, cs]] */
          const double r0 = cs * a0[0] + sn * a1[1], m0 = cs * a0[1] - sn * a1[0];
          const double r1 = sn * a0[1] + cs * a1[0], m1 = cs * a1[1] - sn * a0[0];
          a0[0] = r0; a0[1] = m0; a1[0] = r1; a1[1] = m1;
        }
      } else {
        for (long i = 0; i < S; ++i) {
          if ((i & bw) || !(i & bc)) continue;
          double* a0 = st + 2 * i; double* a1 = st + 2 * (i | bw);
          const double t0 = a0[0], t1 = a0[1];
          a0[0] = a1[0]; a0[1] = a1[1]; a1[0] = t0; a1[1] = t1;
        }
      }
      /* QuantumCircuit.depth() stack, count_ops, idle_wires */
      int lvl = (meta[w] > meta[c] ? meta[w] : meta[c]) + 1;
      if (lvl > 255) lvl = 255;
      meta[w] = meta[c] = (uint8_t)lvl;
      if (ops < 0xffff) ops++;
      used |= (uint16_t)((1u << w) | (1u << c));
      memcpy(meta + 12, &ops, 2); memcpy(meta + 14, &used, 2);
    }
    int depth = 0, nused = 0;
    for (int q = 0; q < 12; ++q) if (meta[q] > depth) depth = meta[q];
    for (int q = 0; q < 16; ++q) nused += (used >> q) & 1;
    const int trunc = status == 0 && (depth >= b->max_depth || ops >= b->max_depth * eta * 2);  /* env.py:129 */

    /* ---- env.py:106-114 */
    double metric_raw;
    if (!uc) {
      double re = 0, im = 0;                                    /* np.vdot(psi, target) */
      for (long i = 0; i < S; ++i) {
        re += st[2 * i] * tg[2 * i] + st[2 * i + 1] * tg[2 * i + 1];
        im += st[2 * i] * tg[2 * i + 1] - st[2 * i + 1] * tg[2 * i];
      }
      const double h = hypot(re, im);
      metric_raw = h * h;
    } else {
      double sr = 0, si = 0;                                    /* np.linalg.norm: real.real + imag.imag */
      for (long i = 0; i < S; ++i) {
        const double dr = tg[2 * i] - st[2 * i], di = tg[2 * i + 1] - st[2 * i + 1];
        sr += dr * dr; si += di * di;
      }
      metric_raw = 1 - 2 * atan(sqrt(sr + si)) / M_PI;
    }
    const double x = (double)depth - b->depth_free;
    const double cost_raw = (x > 0 ? x : 0.0) / b->cost_denom;

    double m = metric_raw, cc = cost_raw, r = 0.0;
    const int sparse = (b->flags & F_SPARSE) != 0;
    const int done = status == 0 && (term || trunc);
    if (status == 0) {
      if (!sparse) {                                            /* env.py:102-104 */
        m = metric_raw - b->last[2 * e]; cc = cost_raw - b->last[2 * e + 1];
        b->last[2 * e] = metric_raw; b->last[2 * e + 1] = cost_raw;
      }
      double rr = m, rc = cc;
      if (sparse && !done) { rr = 0; rc = 0; }                  /* env.py:134 */
      if (b->flags & F_PUNISH) rr -= rc;                        /* env.py:135 */
      r = rr;
      if (b->ep_return) { b->ep_return[e] += r; b->ep_length[e] += 1; }
    }
    b->reward[e] = r; b->metric[e] = m; b->cost[e] = cc;
    b->depth[e] = depth; b->operations[e] = ops; b->used_wires[e] = nused;
    b->terminated[e] = (uint8_t)term; b->truncated[e] = (uint8_t)trunc; b->status[e] = (int8_t)status;

    if (done && b->episode_return) { b->episode_return[e] = b->ep_return[e]; b->episode_length[e] = b->ep_length[e]; }
    if (b->stats) {     /* same fields, same order as QCD_STAT_* in include/qcd_b200.h */
      double* s = b->stats;
      if (status != 0) s[11] += 1;
      else {
        s[10] += 1;
        if (done) {
          s[0] += 1;
          if (b->ep_return) { s[1] += b->ep_return[e]; s[2] += b->ep_length[e]; }
          s[3] += depth; s[4] += ops; s[5] += nused; s[6] += m; s[7] += cc;
          if (trunc) s[9] += 1; else s[8] += 1;
        }
      }
    }
    if (done && (b->flags & F_AUTORESET)) {
      if (b->final_obs) write_obs(b->final_obs + e * b->obs_stride, st, S);
      reset_state(st, S, uc, eta);
      memset(meta, 0, 16);
      if (b->last) { b->last[2 * e] = 0; b->last[2 * e + 1] = 0; }
      if (b->ep_return) { b->ep_return[e] = 0; b->ep_length[e] = 0; }
    }
    if (b->obs) write_obs(b->obs + e * b->obs_stride, st, S);
  }
  return 0;
}

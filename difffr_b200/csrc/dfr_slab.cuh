// Slab domain decomposition of one large scene over several GPUs (BASELINE.json configs[4]).
// The reference has nothing like it (one process, OpenMP); what must be preserved is the result of the
// single-domain step, up to summation order.
//
// Decomposition: the cell index is x-fastest / z-slowest, so a range of z cell layers is a contiguous range of
// the cell-sorted particle arrays.  Rank r owns the layers [own_zlo, own_zhi) of its local grid; one support
// radius (`reach` layers) on each side is the boundary layer it exports / the ghost layer it imports:
//
//     sorted arrays of rank r:   [ ghost_lo | bl_lo ... interior ... bl_hi | ghost_hi ]
//                                            ^own_begin                    ^own_end
//
// Both neighbours sort the shared particles by (cell, particle id), so rank r's bl_hi range and rank r+1's
// ghost_lo range hold the same particles in the same order: a ghost update of any per-particle array is one
// contiguous send + one contiguous receive per side (NCCL over NVLink), no packing.
// Per step (dfr_api.cu: slab_* functions): particles whose new position lies in an export layer - including the
// ones that just crossed the plane - are sent with their full state; the receiver owns what falls into its layers
// and keeps the rest as ghosts; the sender keeps its emigrants as ghosts for this step.  Then every kernel that
// produces a gathered array (x|rho, x|k, v, n|rho) is followed by a ghost update of that array, residual sums, the
// CFL maximum and the per-body force/torque/Jacobian rows are all-reduced, and the rigid bodies (replicated on every
// rank) are advanced identically everywhere.
#pragma once
#include "dfr_kernels.cuh"

namespace dfr {

struct SlabGeom {
  int own_zlo, own_zhi;  // owned z layers (local grid)
  int has_lo, has_hi;    // neighbours exist
  int reach;
};

// Particles of the owned range whose CURRENT position lies in an export layer are appended to the send buffers
// (order irrelevant: the receiver sorts).  misc = (kappa, kappa_v, id, state).
__global__ void k_slab_select(const __grid_constant__ Params P, const SlabGeom G, const StepState *st, const double4 *pos, const double4 *vel,
                              const double *kappa, const double *kappav, const int *pid, const int *pstate, int capacity,
                              double4 *out_pos_lo, double4 *out_vel_lo, double4 *out_misc_lo, double4 *out_pos_hi, double4 *out_vel_hi,
                              double4 *out_misc_hi, int *counts /* [2], zeroed */, int *error_flags) {
  const int i = st->own_begin + blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= st->own_end) return;
  const double4 p = pos[i];
  int cx, cy, cz;
  cell_of(P.grid, p.x, p.y, p.z, cx, cy, cz);
  const bool lo = G.has_lo && cz < G.own_zlo + G.reach;
  const bool hi = G.has_hi && cz >= G.own_zhi - G.reach;
  if (!lo && !hi) return;
  const double4 m = make_double4(kappa[i], kappav[i], __longlong_as_double((long long)pid[i]), __longlong_as_double((long long)pstate[i]));
  if (lo) {
    const int k = atomicAdd(&counts[0], 1);
    if (k < capacity) {
      out_pos_lo[k] = p;
      out_vel_lo[k] = vel[i];
      out_misc_lo[k] = m;
    } else
      atomicOr(error_flags, 8);
  }
  if (hi) {
    const int k = atomicAdd(&counts[1], 1);
    if (k < capacity) {
      out_pos_hi[k] = p;
      out_vel_hi[k] = vel[i];
      out_misc_hi[k] = m;
    } else
      atomicOr(error_flags, 8);
  }
}

// received (kappa, kappa_v, id, state) records -> the persistent arrays, appended at `at`
__global__ void k_slab_unpack(const double4 *misc, int n, int at, double *kappa, double *kappav, int *pid, int *pstate) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const double4 m = misc[k];
  kappa[at + k] = m.x;
  kappav[at + k] = m.y;
  pid[at + k] = (int)__double_as_longlong(m.z);
  pstate[at + k] = (int)__double_as_longlong(m.w);
}

__global__ void k_slab_set_nf(StepState *st, int nf) {
  if (threadIdx.x == 0 && blockIdx.x == 0) st->nf = nf;
}

// cells are sorted by source slot (k_bin_sort_cells); here by particle id, so that two ranks holding the same particles
// in a cell order them identically
// (driven from the particles like k_bin_sort_cells_by_particle: the particle that drew rank 0 sorts its cell)
__global__ void k_bin_sort_cells_by_id(const int *n_ptr, int n_fixed, const int *base_ptr, const unsigned int *cell_start,
                                       const int *cell_of_particle, const int *rank_in_cell, int *sorted_src, const int *pid_src) {
  const int n = n_ptr ? *n_ptr : n_fixed;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n || rank_in_cell[t] != 0) return;
  if (base_ptr) pid_src += *base_ptr;
  const int c = cell_of_particle[t];
  const int s = (int)cell_start[c], e = (int)cell_start[c + 1];
  int *a = sorted_src + s;
  for (int i = 1; i < e - s; i++) {  // insertion sort: cells hold a handful of particles
    const int v = a[i];
    const int kv = pid_src[v];
    int j = i - 1;
    while (j >= 0 && pid_src[a[j]] > kv) {
      a[j + 1] = a[j];
      j--;
    }
    a[j + 1] = v;
  }
}

// owned / boundary-layer ranges of the freshly sorted arrays, from the cell table
__global__ void k_slab_ranges(const __grid_constant__ Params P, const SlabGeom G, StepState *st, const unsigned int *cell_start) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const int layer = P.grid.nx * P.grid.ny;
  const int nf = st->nf;
  const int own_begin = G.has_lo ? (int)cell_start[(size_t)G.own_zlo * layer] : 0;
  const int own_end = G.has_hi ? (int)cell_start[(size_t)G.own_zhi * layer] : nf;
  st->own_begin = own_begin;
  st->own_end = own_end;
  st->slab_ranges[0] = own_begin;
  st->slab_ranges[1] = own_end;
  st->slab_ranges[2] = G.has_lo ? (int)cell_start[(size_t)(G.own_zlo + G.reach) * layer] : own_begin;  // end of bl_lo
  st->slab_ranges[3] = G.has_hi ? (int)cell_start[(size_t)(G.own_zhi - G.reach) * layer] : own_end;    // begin of bl_hi
  st->slab_ranges[4] = nf;
  st->slab_ranges[5] = st->slab_ranges[2] - own_begin;  // particles in my low / high boundary layer
  st->slab_ranges[6] = own_end - st->slab_ranges[3];
  st->slab_ranges[7] = own_end;  // sent together with [6]: the high neighbour mirrors its low boundary layer behind my owned range
  // particles that escaped beyond the ghost layers cannot be handled (they would need more than one hop)
  const int ghost_lo_begin = G.has_lo ? (int)cell_start[(size_t)(G.own_zlo - G.reach) * layer] : 0;
  const int ghost_hi_end = G.has_hi ? (int)cell_start[(size_t)(G.own_zhi + G.reach) * layer] : nf;
  if (ghost_lo_begin != 0 || ghost_hi_end != nf) atomicOr(&st->error_flags, 16);
}

// TimeStepDiffDFSPH::pressureSolve / divergenceSolve stopping rules (:711-743, :828-861) on the all-reduced residual
// (NCCL transport: the sum over the slabs arrives in st->res_sum)
template <bool PRESSURE>
__global__ void k_solver_decide(const __grid_constant__ Params P, StepState *st) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  if (!(PRESSURE ? st->prs_active : st->div_active)) return;
  solver_decide<PRESSURE>(P, st, st->res_sum / (double)P.n_global, 0ull, 0);
}

// ---------------------------------------------------------------------------------------------
// Small all-reduces over peer memory.  Every rank maps every rank's mailbox (cudaIpc) and writes its contribution
// straight into all of them over NVLink; a rank then finds the n contributions in its own mailbox and combines them in
// rank order, so that every rank computes the same bits.  One block, no collective library, no host: the call sits in
// the stream (or in a recorded step graph) like any other kernel.  Sequence numbers make the two payload slots safe: a
// rank can only write its contribution to all-reduce q + 2 into the slot of q after it has finished q + 1, which needs
// every peer's contribution to q + 1, which a peer only sends after it has read q.
// ---------------------------------------------------------------------------------------------
#define DFR_MAX_SLABS 16
struct SlabMail {
  double *box[DFR_MAX_SLABS];               // rank r's mailbox: [2 slots][n ranks][cap] doubles
  unsigned long long *flag[DFR_MAX_SLABS];  // rank r's flags: [n ranks], the sequence number rank k has delivered
  unsigned long long *seq;                  // my count of all-reduces so far; never reset (the peers' flags are not either)
  int n, rank, cap;
};
enum { MAIL_SUM_F64 = 0, MAIL_MAX_U64 = 1 };
// data[0..count) <- combination over all ranks; called by every thread of one block
template <int OP>
__device__ __forceinline__ void mailbox_allreduce(const SlabMail &M, StepState *st, double *data, int count, unsigned long long timeout_ns) {
  __shared__ unsigned long long seq_s;
  __shared__ int fail_s;
  if (threadIdx.x == 0) {
    seq_s = ++(*M.seq);
    fail_s = (st->error_flags & 32) ? 1 : 0;  // a peer already failed to answer once: do not stack further time-outs
  }
  __syncthreads();
  const unsigned long long q = seq_s;
  const size_t slot = (size_t)(q & 1ull) * (size_t)M.n * (size_t)M.cap;
  if (count > M.cap) count = M.cap;
  for (int p = 0; p < M.n; p++) {
    double *dst = M.box[p] + slot + (size_t)M.rank * (size_t)M.cap;
    for (int k = threadIdx.x; k < count; k += blockDim.x) dst[k] = data[k];
  }
  __threadfence_system();
  __syncthreads();
  if ((int)threadIdx.x < M.n) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(M.flag[threadIdx.x] + M.rank), "l"(q) : "memory");
    unsigned long long t0, t1, v;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    const unsigned long long *mine = M.flag[M.rank] + threadIdx.x;
    while (!fail_s) {
      asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(mine) : "memory");
      if (v >= q) break;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
      if (t1 - t0 > timeout_ns) {
        fail_s = 1;
        break;
      }
    }
  }
  __syncthreads();
  if (fail_s) {
    if (threadIdx.x == 0) atomicOr(&st->error_flags, 32);
    return;
  }
  const double *in = M.box[M.rank] + slot;
  for (int k = threadIdx.x; k < count; k += blockDim.x) {
    if (OP == MAIL_SUM_F64) {
      double s = 0.0;
      for (int r = 0; r < M.n; r++) s += __ldcv(in + (size_t)r * M.cap + k);  // volatile loads: the peers wrote these lines
      data[k] = s;
    } else {
      unsigned long long m = 0ull;
      for (int r = 0; r < M.n; r++) m = max(m, (unsigned long long)__double_as_longlong(__ldcv(in + (size_t)r * M.cap + k)));
      data[k] = __longlong_as_double((long long)m);
    }
  }
  __syncthreads();
}
#define MAIL_TIMEOUT_NS 5000000000ull
// the residual of a Jacobi iteration summed over the slabs + the stopping rule (replaces ncclAllReduce + k_solver_decide)
template <bool PRESSURE>
__global__ void __launch_bounds__(64) k_slab_residual_decide(const __grid_constant__ Params P, StepState *st, const SlabMail M,
                                                             unsigned long long cond, int policy) {
  if (!(PRESSURE ? st->prs_active : st->div_active)) {
    if (cond && threadIdx.x == 0) cudaGraphSetConditional((cudaGraphConditionalHandle)cond, 0u);
    return;
  }
  mailbox_allreduce<MAIL_SUM_F64>(M, st, &st->res_sum, 1, MAIL_TIMEOUT_NS);
  if (threadIdx.x != 0) return;
  if (st->error_flags & 32) {  // a peer did not answer: close the solve, the host reports the error
    if (PRESSURE) st->prs_active = 0; else st->div_active = 0;
    if (cond) cudaGraphSetConditional((cudaGraphConditionalHandle)cond, 0u);
    return;
  }
  solver_decide<PRESSURE>(P, st, st->res_sum / (double)P.n_global, cond, policy);
}
// max |v + a h|^2 over the slabs (ordered bits of positive doubles)
__global__ void __launch_bounds__(64) k_slab_max_u64(StepState *st, const SlabMail M, unsigned long long *value) {
  mailbox_allreduce<MAIL_MAX_U64>(M, st, reinterpret_cast<double *>(value), 1, MAIL_TIMEOUT_NS);
}
// the per-body force / torque / Jacobian rows summed over the slabs
__global__ void __launch_bounds__(256) k_slab_sum_f64(StepState *st, const SlabMail M, double *buf, int count) {
  mailbox_allreduce<MAIL_SUM_F64>(M, st, buf, count, MAIL_TIMEOUT_NS);
}

// Peer-memory transport: "my boundary rows of this pass are in your ghost range" / "are yours in mine?".
// One thread: release my writes at system scope, raise the counter in both neighbours' flag words, then spin on mine.
// flags[0] is written by the low neighbour, flags[1] by the high one.  A wait that lasts longer than `timeout_ns` gives
// up and raises error bit 32 (a peer that failed must not hang this GPU).
__device__ __forceinline__ bool slab_signal_wait(unsigned long long *peer_lo_flag, unsigned long long *peer_hi_flag,
                                                 volatile unsigned long long *my_flags, unsigned long long *pass_counter,
                                                 unsigned long long timeout_ns, int *error_flags) {
  // the pass number lives on the device (both neighbours make the same sequence of calls; never reset, like the flags),
  // so that the kernel can be part of a recorded step graph
  const unsigned long long value = ++(*pass_counter);
  if (*error_flags & 32) return false;  // a neighbour already failed to answer once: do not stack further time-outs
  __threadfence_system();
  if (peer_lo_flag) asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(peer_lo_flag), "l"(value) : "memory");
  if (peer_hi_flag) asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(peer_hi_flag), "l"(value) : "memory");
  unsigned long long t0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  for (int side = 0; side < 2; side++) {
    if (!(side == 0 ? peer_lo_flag : peer_hi_flag)) continue;
    for (;;) {
      unsigned long long v;
      asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"((const unsigned long long *)(my_flags + side)) : "memory");
      if (v >= value) break;
      unsigned long long t1;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
      if (t1 - t0 > timeout_ns) {
        atomicOr(error_flags, 32);
        return false;
      }
    }
  }
  return true;
}
__global__ void k_slab_signal_wait(unsigned long long *peer_lo_flag, unsigned long long *peer_hi_flag, volatile unsigned long long *my_flags,
                                   unsigned long long *pass_counter, unsigned long long timeout_ns, int *error_flags) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  slab_signal_wait(peer_lo_flag, peer_hi_flag, my_flags, pass_counter, timeout_ns, error_flags);
}

// ---------------------------------------------------------------------------------------------
// Device-side particle exchange (the head of a replayed slab step: no NCCL call, no host read-back).
// Every rank exports one receive area (cudaIpc): a header of ints and, per side, room for `cap` (pos, vel, misc)
// records.  k_slab_select stores the particles of my export layers straight into the neighbours' areas over NVLink,
// k_slab_xchg_post_rows tells them how many and waits for theirs (same flag protocol as the ghost updates),
// k_slab_xchg_unpack appends what arrived behind my owned range; after the re-sort k_slab_xchg_post_layers exchanges
// the sizes of the shared layers (the neighbours' ghost stores need my own_end; a mismatch is an error, not a hang).
// A neighbour can only write its next step's rows after all of this step's ghost-update passes, which I answer after
// my unpack: one area per side is enough.
// ---------------------------------------------------------------------------------------------
enum { XH_ROWS_LO = 0, XH_ROWS_HI = 1, XH_LO_BL = 2, XH_LO_OWN_END = 3, XH_HI_BL = 4, XH_INTS = 64 };
struct SlabXchg {
  char *mine, *peer_lo, *peer_hi;  // receive areas (peer_*: null without that neighbour)
  int cap;
};
__host__ __device__ __forceinline__ int *xchg_hdr(char *area) { return reinterpret_cast<int *>(area); }
// rows sent by the neighbour on `side` (0: low, 1: high) of the area's owner; arr 0 pos, 1 vel, 2 misc
__host__ __device__ __forceinline__ double4 *xchg_rows(char *area, int side, int arr, int cap) {
  return reinterpret_cast<double4 *>(area + XH_INTS * sizeof(int)) + (size_t)(side * 3 + arr) * (size_t)cap;
}
__host__ __device__ __forceinline__ size_t xchg_bytes(int cap) { return XH_INTS * sizeof(int) + (size_t)6 * (size_t)cap * sizeof(double4); }

__device__ __forceinline__ double4 ldcv4(const double4 *p) {  // the neighbour wrote these rows: volatile loads
  const double2 a = __ldcv(reinterpret_cast<const double2 *>(p)), b = __ldcv(reinterpret_cast<const double2 *>(p) + 1);
  return make_double4(a.x, a.y, b.x, b.y);
}
__global__ void k_slab_xchg_post_rows(const int *counts, const SlabXchg X, unsigned long long *peer_lo_flag, unsigned long long *peer_hi_flag,
                                      volatile unsigned long long *my_flags, unsigned long long *pass_counter, unsigned long long timeout_ns,
                                      int *error_flags) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  if (X.peer_lo) xchg_hdr(X.peer_lo)[XH_ROWS_HI] = min(counts[0], X.cap);  // I am its high neighbour
  if (X.peer_hi) xchg_hdr(X.peer_hi)[XH_ROWS_LO] = min(counts[1], X.cap);
  slab_signal_wait(peer_lo_flag, peer_hi_flag, my_flags, pass_counter, timeout_ns, error_flags);
}
__global__ void k_slab_xchg_unpack(StepState *st, const SlabXchg X, int nf_cap, double4 *pos, double4 *vel, double *kappa, double *kappav,
                                   int *pid, int *pstate) {
  const int *h = xchg_hdr(X.mine);
  const int at = st->own_end;
  int n_lo = X.peer_lo ? min(max(__ldcv(h + XH_ROWS_LO), 0), X.cap) : 0;
  int n_hi = X.peer_hi ? min(max(__ldcv(h + XH_ROWS_HI), 0), X.cap) : 0;
  if (st->error_flags & 32) n_lo = n_hi = 0;  // a neighbour did not answer: the header is stale
  const int room = max(nf_cap - at, 0);
  const bool over = n_lo + n_hi > room;
  if (over) {
    n_lo = min(n_lo, room);
    n_hi = min(n_hi, room - n_lo);
  }
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k == 0) {
    st->nf = (at - st->own_begin) + n_lo + n_hi;  // what the re-sort works on: [own_begin, own_end + received)
    if (over) atomicOr(&st->error_flags, 8);
  }
  if (k >= n_lo + n_hi) return;
  const int side = k < n_lo ? 0 : 1, r = k < n_lo ? k : k - n_lo;
  const double4 m = ldcv4(xchg_rows(X.mine, side, 2, X.cap) + r);
  pos[at + k] = ldcv4(xchg_rows(X.mine, side, 0, X.cap) + r);
  vel[at + k] = ldcv4(xchg_rows(X.mine, side, 1, X.cap) + r);
  kappa[at + k] = m.x;
  kappav[at + k] = m.y;
  pid[at + k] = (int)__double_as_longlong(m.z);
  pstate[at + k] = (int)__double_as_longlong(m.w);
}
// k_bin_count over [own_begin, own_begin + nf) with the range read on the device
__global__ void k_slab_bin_count(const __grid_constant__ Params P, const StepState *st, const double4 *pos, unsigned int *cell_count,
                                 int *cell_of_particle, int *rank_in_cell) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= st->nf) return;
  const double4 p = pos[st->own_begin + i];
  int cx, cy, cz;
  cell_of(P.grid, p.x, p.y, p.z, cx, cy, cz);
  const int c = cell_lin(P.grid, cx, cy, cz);
  cell_of_particle[i] = c;
  rank_in_cell[i] = (int)atomicAdd(&cell_count[c], 1u);
}
__global__ void k_slab_xchg_post_layers(StepState *st, const SlabXchg X, unsigned long long *peer_lo_flag, unsigned long long *peer_hi_flag,
                                        volatile unsigned long long *my_flags, unsigned long long *pass_counter,
                                        unsigned long long timeout_ns) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  if (X.peer_lo) xchg_hdr(X.peer_lo)[XH_HI_BL] = st->slab_ranges[5];
  if (X.peer_hi) {
    xchg_hdr(X.peer_hi)[XH_LO_BL] = st->slab_ranges[6];
    xchg_hdr(X.peer_hi)[XH_LO_OWN_END] = st->slab_ranges[7];
  }
  const bool ok = slab_signal_wait(peer_lo_flag, peer_hi_flag, my_flags, pass_counter, timeout_ns, &st->error_flags);
  const int *h = xchg_hdr(X.mine);
  const int ghost_lo = st->slab_ranges[0], ghost_hi = st->slab_ranges[4] - st->slab_ranges[1];
  if (!ok || (X.peer_lo && __ldcv(h + XH_LO_BL) != ghost_lo) || (X.peer_hi && __ldcv(h + XH_HI_BL) != ghost_hi)) {
    // the two sides of a plane disagree about the shared layers: no ghost row may be stored (the ranges would not match)
    if (ok) atomicOr(&st->error_flags, 64);
    st->slab_ranges[2] = st->slab_ranges[0];
    st->slab_ranges[3] = st->slab_ranges[1];
  }
}

// k_body_reduce split in two around the all-reduce of the per-body rows
__global__ void k_body_rows_to_buf(const BodyDev *bodies, double *acc_rows, double *buf) {
  const BodyDev &B = bodies[blockIdx.x];
  __shared__ double tot[ACC_N];
  __shared__ double part[BR_SLICES][ACC_N];
  body_rows_sum(B, acc_rows, part, tot);
  for (int k = threadIdx.x; k < ACC_N; k += blockDim.x) buf[(size_t)blockIdx.x * ACC_N + k] = tot[k];
}
__global__ void k_body_buf_apply(BodyDev *bodies, const double *buf) {
  BodyDev &B = bodies[blockIdx.x];
  if (!B.dynamic || threadIdx.x != 0) return;
  const double *tot = buf + (size_t)blockIdx.x * ACC_N;
  B.force += mk3(tot[ACC_F], tot[ACC_F + 1], tot[ACC_F + 2]);
  B.torque += mk3(tot[ACC_T], tot[ACC_T + 1], tot[ACC_T + 2]);
  if (!B.animated) {
    for (int k = 0; k < 9; k++) {
      B.net_f_v.a[k] = tot[ACC_FV + k];
      B.net_f_x.a[k] = tot[ACC_FX + k];
      B.net_f_w.a[k] = tot[ACC_FW + k];
      B.net_t_v.a[k] = tot[ACC_TV + k];
      B.net_t_x.a[k] = tot[ACC_TX + k];
      B.net_t_w.a[k] = tot[ACC_TW + k];
    }
    for (int k = 0; k < 12; k++) {
      B.net_f_q.a[k] = tot[ACC_FQ + k];
      B.net_t_q.a[k] = tot[ACC_TQ + k];
    }
  }
}

}  // namespace dfr

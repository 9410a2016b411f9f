// Hand-written sm_100a kernels of the differentiable DFSPH time step.
// Each kernel cites the reference loop it replaces (paths relative to the reference checkout).
// Layout rules: fluid particles are re-sorted into cell order every step, positions
// and velocities are 32-byte double4 records (one sector per gathered neighbour), neighbour
// lists are warp-interleaved so index loads coalesce, no tensor cores (no dense contraction).
#pragma once
#include "dfr_types.cuh"

namespace dfr {

#define DFR_EPS 1.0e-5  // m_eps, TimeStepDiffDFSPH.h:26
// gathers in flight per record array and neighbour batch (tuned on B200, see profiles/)
#ifndef DFR_RHO_U
#define DFR_RHO_U 1
#endif
#ifndef DFR_PUSH_U
#define DFR_PUSH_U 2
#endif
#ifndef DFR_NP_U
#define DFR_NP_U 1
#endif
#ifndef DFR_DF_U
#define DFR_DF_U 4
#endif
#ifndef DFR_RHO_BLOCKS
#define DFR_RHO_BLOCKS 8
#endif
// resident 128-thread blocks per SM asked of the compiler (register cap = 512 / blocks) for the fused k_rho variants
// and for k_push
#ifndef DFR_RHOX_BLOCKS
#define DFR_RHOX_BLOCKS 7
#endif
#ifndef DFR_RHONP_BLOCKS
#define DFR_RHONP_BLOCKS 5
#endif
#ifndef DFR_PUSH_BLOCKS
#define DFR_PUSH_BLOCKS 8
#endif
struct Rec2 {
  double4 a, b;
};
struct Rec3 {
  double4 a, b, c;
};
#define DFR_FULL 0xffffffffu

// ---------------------------------------------------------------------------------------------
// SM-local block scheduling for the neighbour-gather kernels.  A plain grid hands consecutive 128-particle blocks
// to different SMs, so the ~8 blocks resident on one SM gather from 8 unrelated neighbourhoods and thrash its L1
// (profiles/r1d: 60 % L1 hit rate although a block re-uses every line ~16 times).  Here the grid is persistent
// (SMs x resident blocks) and every SM owns one contiguous range of virtual blocks: its resident CTAs claim
// consecutive blocks of that range, i.e. spatially adjacent particles whose neighbourhoods overlap.  A CTA that runs
// out of work sweeps the other SMs' ranges, so every block is processed no matter how CTAs were placed.
// Two counter sets alternate between launches; each launch clears the set the next one will use.
// MEASURED (profiles/r1e_sm_local_experiment.md, B200, 1 M particles): L1 hit rate 62 -> 77 %, L2 reads -30 %, but the
// step got SLOWER (3.81 -> 5.09 ms): the gather kernels are bound by L1 wavefronts per gather (~16 lines per warp-wide
// gather), not by the hit rate, and the per-block barriers of a persistent CTA add stalls.  Kept as a build option
// (-DDFR_SM_LOCAL=1), off by default.
// ---------------------------------------------------------------------------------------------
#ifndef DFR_SM_LOCAL
#define DFR_SM_LOCAL 0
#endif
// Virtual 128-particle blocks per CTA of the list kernels: a CTA of 128 * DFR_CTA_VB threads works on that many
// CONSECUTIVE virtual blocks, i.e. on one contiguous run of cell-sorted particles whose neighbourhoods overlap, so that
// the warps sharing an SM's L1 at any time gather from the same few hundred KB instead of from 8 unrelated strips
// (tools/gather_locality.py: the gathers of 128 consecutive particles can hit L1 at most 70 % of the time on the bench
// scene's 93-cell rows, those of 1024 consecutive particles 85 %).  Indexing inside a virtual block is unchanged, so every
// sum has the same order for every DFR_CTA_VB.
#ifndef DFR_CTA_VB
#define DFR_CTA_VB 1
#endif
#if DFR_SM_LOCAL && DFR_CTA_VB != 1
#error "DFR_SM_LOCAL and DFR_CTA_VB are alternatives"
#endif
#define DFR_CTA_THREADS (128 * DFR_CTA_VB)
#define DFR_TID (threadIdx.x & 127)
#define DFR_RESIDENT(blocks) ((blocks) / DFR_CTA_VB > 0 ? (blocks) / DFR_CTA_VB : 1)
#define DFR_SCHED_STRIDE 256
struct VSched {
  unsigned int *ctr;  // [2][DFR_SCHED_STRIDE]
  int parity, nsm, per, nvb;
};
__device__ __forceinline__ void vsched_prologue(const VSched &S) {
#if DFR_SM_LOCAL
  if (blockIdx.x == 0)
    for (int t = threadIdx.x; t < DFR_SCHED_STRIDE; t += blockDim.x) S.ctr[(1 - S.parity) * DFR_SCHED_STRIDE + t] = 0u;
#endif
}
struct VState {
  int sm, tried;
};
__device__ __forceinline__ VState vsched_begin(const VSched &S) {
  VState v;
  unsigned int smid;
  asm("mov.u32 %0, %%smid;" : "=r"(smid));
  v.sm = (int)(smid % (unsigned int)S.nsm);
  v.tried = 0;
  return v;
}
__device__ __forceinline__ int vsched_next(const VSched &S, VState &v, int *sh) {
  __syncthreads();
  if (threadIdx.x == 0) {
    int vb = -1;
    while (v.tried < S.nsm) {
      const unsigned int k = atomicAdd(&S.ctr[S.parity * DFR_SCHED_STRIDE + v.sm], 1u);
      const long long cand = (long long)v.sm * S.per + k;
      if (k < (unsigned int)S.per && cand < S.nvb) {
        vb = (int)cand;
        break;
      }
      v.sm = (v.sm + 1 == S.nsm) ? 0 : v.sm + 1;
      v.tried++;
    }
    *sh = vb;
  }
  __syncthreads();
  return *sh;
}
#if DFR_SM_LOCAL
#define DFR_VB_LOOP(S)               \
  __shared__ int vb_sh_;             \
  VState vb_state_ = vsched_begin(S); \
  for (int vb_ = vsched_next(S, vb_state_, &vb_sh_); vb_ >= 0; vb_ = vsched_next(S, vb_state_, &vb_sh_))
#else
#define DFR_VB_LOOP(S) for (int vb_ = blockIdx.x * DFR_CTA_VB + (threadIdx.x >> 7), once_ = 1; once_; once_ = 0)
#endif

// read-only 32-byte record load as ONE 256-bit instruction (sm_100: LDG.E.ENL2.256.CONSTANT).  A gathered
// record then costs one L1 tag/data wavefront per distinct 128-byte line instead of two (profiles/r1a: the
// neighbour passes are bound by L1 data-stage wavefronts, not by HBM or the FP64 pipe).
__device__ __forceinline__ double4 ldg4(const double4 *p) {
  double4 r;
  asm("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(r.x), "=d"(r.y), "=d"(r.z), "=d"(r.w) : "l"(p));
  return r;
}
// 256-bit store of one record (STG.E.ENL2.256)
__device__ __forceinline__ void stg4(double4 *p, const double4 &v) {
  asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(v.x), "d"(v.y), "d"(v.z), "d"(v.w) : "memory");
}
// plain (coherent) 256-bit load of a record this kernel family also writes
__device__ __forceinline__ double4 ld4(const double4 *p) {
  double4 r;
  asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(r.x), "=d"(r.y), "=d"(r.z), "=d"(r.w) : "l"(p) : "memory");
  return r;
}

// ---------------------------------------------------------------------------------------------
// Slab decomposition, peer-memory transport (dfr_slab.cuh): the kernel that produces a gathered array also writes the
// rows of its boundary layers straight into the neighbour GPU's ghost range (stores over NVLink to a cudaIpc-mapped
// peer buffer), so that a ghost update costs no extra pass over the data and no collective launch - only a flag.
// lo / hi are null on a single context or with the NCCL transport.
// ---------------------------------------------------------------------------------------------
// The row ranges come from device memory (StepState::slab_ranges, written by k_slab_ranges after every exchange; the low
// neighbour's own_end arrives in the exchange's count message), so that a recorded step graph stays valid from step to step.
struct GhostOut {
  double4 *lo, *hi;          // the neighbours' arrays (base addresses)
  const int *ranges;         // StepState::slab_ranges: [0] own_begin [1] own_end [2] end of my low boundary layer [3] begin of the high one
  const int *lo_nb_own_end;  // the low neighbour's own_end: its ghost_hi range, which mirrors my low boundary layer, starts there
};
__device__ __forceinline__ void ghost_store(const GhostOut &g, int i, const double4 &v) {
  if (g.lo) {
    const int b = g.ranges[0];
    if (i >= b && i < g.ranges[2]) stg4(g.lo + *g.lo_nb_own_end + (i - b), v);
  }
  if (g.hi) {  // the high neighbour's ghost_lo range starts at 0
    const int b = g.ranges[3];
    if (i >= b && i < g.ranges[1]) stg4(g.hi + (i - b), v);
  }
}

// ---------------------------------------------------------------------------------------------
// cell coordinates
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
__device__ __forceinline__ void cell_of(const GridGeom &g, double x, double y, double z, int &cx, int &cy, int &cz) {
  cx = clampi((int)floor((x - g.ox) * g.inv_cell), 0, g.nx - 1);
  cy = clampi((int)floor((y - g.oy) * g.inv_cell), 0, g.ny - 1);
  cz = clampi((int)floor((z - g.oz) * g.inv_cell) - g.z_shift, 0, g.nz - 1);
}
__device__ __forceinline__ int cell_lin(const GridGeom &g, int cx, int cy, int cz) { return (cz * g.ny + cy) * g.nx + cx; }

// CompactNSearch distance predicate (upstream @3f11ece1): l2 = dx*dx; l2 += dy*dy; l2 += dz*dz;
// l2 < r2.  Written with explicit round-to-nearest intrinsics so that nvcc cannot contract it into
// FMAs: neighbour sets then match the oracle (built with -ffp-contract=off) bit for bit.
__device__ __forceinline__ double dist2_exact(double ax, double ay, double az, double bx, double by, double bz) {
  double t = __dsub_rn(ax, bx);
  double l2 = __dmul_rn(t, t);
  t = __dsub_rn(ay, by);
  l2 = __dadd_rn(l2, __dmul_rn(t, t));
  t = __dsub_rn(az, bz);
  l2 = __dadd_rn(l2, __dmul_rn(t, t));
  return l2;
}

// ---------------------------------------------------------------------------------------------
// SPH kernel functions (SPHKernels.h:37-55, 80-102, 123-150, 483-499, 544-556)
// ---------------------------------------------------------------------------------------------
// Returns W(|r|) and writes c with gradW(r) = c * r.
__device__ __forceinline__ double cubic_W_and_grad(const Params &P, double r2, double &c) {
  double rl, rinv;
  if (r2 > 1.0e-10) {  // rl > 1e-5
    rinv = rsqrt(r2);
    rl = r2 * rinv;
  } else {
    rl = sqrt(r2);
    rinv = 0.0;
  }
  const double q = rl * P.inv_h;
  double w;
  if (q <= 0.5) {
    const double q2 = q * q;
    w = P.k_cubic * (6.0 * q2 * q - 6.0 * q2 + 1.0);
    c = P.l_cubic * q * (3.0 * q - 2.0);
  } else {
    const double f = fmax(1.0 - q, 0.0);
    w = P.k_cubic * 2.0 * f * f * f;
    c = -P.l_cubic * f * f;
  }
  c *= rinv * P.inv_h;
  return w;
}
__device__ __forceinline__ double cubic_grad_coeff(const Params &P, double r2) {
  if (!(r2 > 1.0e-10)) return 0.0;
  const double rinv = rsqrt(r2);
  const double q = r2 * rinv * P.inv_h;
  double c;
  if (q <= 0.5)
    c = P.l_cubic * q * (3.0 * q - 2.0);
  else {
    const double f = fmax(1.0 - q, 0.0);
    c = -P.l_cubic * f * f;
  }
  return c * rinv * P.inv_h;
}
__device__ __forceinline__ d3 cubic_gradW(const Params &P, d3 r) { return cubic_grad_coeff(P, dot(r, r)) * r; }
__device__ __forceinline__ double cubic_W(const Params &P, double r2) {
  const double q = sqrt(r2) * P.inv_h;
  if (q > 1.0) return 0.0;
  if (q <= 0.5) {
    const double q2 = q * q;
    return P.k_cubic * (6.0 * q2 * q - 6.0 * q2 + 1.0);
  }
  const double f = 1.0 - q;
  return P.k_cubic * 2.0 * f * f * f;
}
// gradW and gradGradW of r (CubicKernel::gradGradW, SPHKernels.h:123-150)
__device__ __forceinline__ void cubic_grad_gradgrad(const Params &P, d3 r, d3 &g, m33 &H) {
  const double r2 = dot(r, r);
  if (!(r2 > 1.0e-10)) {
    g = mk3(0, 0, 0);
    H = m33::zero();
    return;
  }
  const double rl = sqrt(r2);
  const double q = rl * P.inv_h;
  if (q > 1.0) {
    g = mk3(0, 0, 0);
    H = m33::zero();
    return;
  }
  const double s = 1.0 / (rl * P.support_radius);
  const d3 gradq = s * r;
  double c1, c2;
  if (q <= 0.5) {
    c1 = P.l_cubic * q * (3.0 * q - 2.0);
    c2 = P.l_cubic * (6.0 * q - 2.0);
  } else {
    const double f = 1.0 - q;
    c1 = P.l_cubic * (-f * f);
    c2 = P.l_cubic * 2.0 * f;
  }
  g = c1 * gradq;
  const m33 rr = outer(r, r);
  const m33 gg = outer(gradq, gradq);
  const double ir2 = 1.0 / r2;
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) H.a[i * 3 + j] = c1 * s * ((i == j ? 1.0 : 0.0) - rr.a[i * 3 + j] * ir2) + c2 * gg.a[i * 3 + j];
}
__device__ __forceinline__ double cube(double x) { return x * x * x; }
__device__ __forceinline__ double cohesion_W(const Params &P, double r2) {
  const double h = P.support_radius;
  if (!(r2 <= h * h)) return 0.0;
  const double r1 = sqrt(r2);
  const double r3 = r2 * r1;
  if (r1 > 0.5 * h) return P.coh_k * cube(h - r1) * r3;
  return P.coh_k * 2.0 * cube(h - r1) * r3 - P.coh_c;
}
__device__ __forceinline__ double adhesion_W(const Params &P, double r2) {
  const double h = P.support_radius;
  if (!(r2 <= h * h)) return 0.0;
  const double rl = sqrt(r2);
  if (rl > 0.5 * h) return P.adh_k * pow(-4.0 * r2 / h + 6.0 * rl - 2.0 * h, 0.25);
  return 0.0;
}

// ---------------------------------------------------------------------------------------------
// generic helpers
// ---------------------------------------------------------------------------------------------
__global__ void k_fill_i32(int *p, int v, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

__global__ void k_iota_i32(int *p, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = (int)i;
}

// exclusive scan of unsigned ints, 3 kernels (block scan, scan of block sums, add offsets)
#define SCAN_THREADS 512
#define SCAN_ITEMS 8
#define SCAN_TILE (SCAN_THREADS * SCAN_ITEMS)
__global__ void k_scan_tiles(const unsigned int *in, unsigned int *out, unsigned int *tile_sums, size_t n) {
  __shared__ unsigned int warp_sums[SCAN_THREADS / 32];
  const size_t base = (size_t)blockIdx.x * SCAN_TILE + (size_t)threadIdx.x * SCAN_ITEMS;
  unsigned int v[SCAN_ITEMS];
  unsigned int sum = 0;
  static_assert(SCAN_ITEMS == 8, "two 128-bit accesses per thread");
  const bool full = (base + SCAN_ITEMS <= n);
  if (full) {
    const uint4 a = *reinterpret_cast<const uint4 *>(in + base), b = *reinterpret_cast<const uint4 *>(in + base + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
    v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  } else {
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) v[k] = (base + k < n) ? in[base + k] : 0u;
  }
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; k++) sum += v[k];
  // inclusive scan of thread sums
  unsigned int incl = sum;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    unsigned int t = __shfl_up_sync(DFR_FULL, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) warp_sums[wid] = incl;
  __syncthreads();
  if (wid == 0) {
    unsigned int ws = (lane < SCAN_THREADS / 32) ? warp_sums[lane] : 0u;
    unsigned int wi = ws;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      unsigned int t = __shfl_up_sync(DFR_FULL, wi, o);
      if (lane >= o) wi += t;
    }
    if (lane < SCAN_THREADS / 32) warp_sums[lane] = wi - ws;  // exclusive
    if (lane == SCAN_THREADS / 32 - 1) tile_sums[blockIdx.x] = wi;
  }
  __syncthreads();
  unsigned int run = warp_sums[wid] + (incl - sum);
  unsigned int o8[SCAN_ITEMS];
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; k++) {
    o8[k] = run;
    run += v[k];
  }
  if (full) {
    *reinterpret_cast<uint4 *>(out + base) = make_uint4(o8[0], o8[1], o8[2], o8[3]);
    *reinterpret_cast<uint4 *>(out + base + 4) = make_uint4(o8[4], o8[5], o8[6], o8[7]);
  } else {
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++)
      if (base + k < n) out[base + k] = o8[k];
  }
}
// single block: exclusive scan of the tile sums in place; writes the grand total to *total
__global__ void k_scan_sums(unsigned int *tile_sums, int ntiles, unsigned int *total) {
  __shared__ unsigned int warp_sums[32];
  __shared__ unsigned int carry_s;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int base = 0; base < ntiles; base += blockDim.x) {
    const int i = base + threadIdx.x;
    const unsigned int v = (i < ntiles) ? tile_sums[i] : 0u;
    unsigned int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      unsigned int t = __shfl_up_sync(DFR_FULL, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) warp_sums[wid] = incl;
    __syncthreads();
    if (wid == 0) {
      unsigned int ws = warp_sums[lane];
      unsigned int wi = ws;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        unsigned int t = __shfl_up_sync(DFR_FULL, wi, o);
        if (lane >= o) wi += t;
      }
      warp_sums[lane] = wi - ws;
    }
    __syncthreads();
    const unsigned int carry = carry_s;
    const unsigned int excl = carry + warp_sums[wid] + (incl - v);
    if (i < ntiles) tile_sums[i] = excl;
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) carry_s = excl + v;
    __syncthreads();
  }
  if (threadIdx.x == 0 && total) *total = carry_s;
}
// four elements per thread (SCAN_TILE is a multiple of 4, so the four share one tile)
__global__ void k_scan_add(unsigned int *out, const unsigned int *tile_sums, size_t n) {
  const size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i >= n) return;
  const unsigned int add = tile_sums[i / SCAN_TILE];
  if (i + 4 <= n) {
    uint4 v = *reinterpret_cast<uint4 *>(out + i);
    v.x += add; v.y += add; v.z += add; v.w += add;
    *reinterpret_cast<uint4 *>(out + i) = v;
  } else {
    for (size_t k = i; k < n; k++) out[k] += add;
  }
}

// ---------------------------------------------------------------------------------------------
// cell binning / counting sort (replaces CompactNSearch's hash-grid build; call site
// Simulation.cpp:746 find_neighbors, and z_sort Simulation.cpp:755)
// ---------------------------------------------------------------------------------------------
__global__ void k_bin_count(const __grid_constant__ Params P, const double4 *pos, const int *n_ptr, int n_fixed, unsigned int *cell_count,
                            int *cell_of_particle, int *rank_in_cell) {
  const int n = n_ptr ? *n_ptr : n_fixed;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double4 p = pos[i];
  int cx, cy, cz;
  cell_of(P.grid, p.x, p.y, p.z, cx, cy, cz);
  const int c = cell_lin(P.grid, cx, cy, cz);
  cell_of_particle[i] = c;
  rank_in_cell[i] = (int)atomicAdd(&cell_count[c], 1u);
}
// slot[cell_start + rank] = i, then each cell's slice is sorted by i so the permutation is the
// (deterministic) stable counting sort irrespective of the order the atomics were served in.
__global__ void k_bin_scatter(const int *n_ptr, int n_fixed, const unsigned int *cell_start, const int *cell_of_particle,
                              const int *rank_in_cell, int *sorted_src) {
  const int n = n_ptr ? *n_ptr : n_fixed;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  sorted_src[cell_start[cell_of_particle[i]] + rank_in_cell[i]] = i;
}
__global__ void k_bin_sort_cells(const unsigned int *cell_start, int ncells, int *sorted_src) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ncells) return;
  const int s = (int)cell_start[c], e = (int)cell_start[c + 1];
  const int m = e - s;
  if (m < 2) return;
  int *a = sorted_src + s;
  if (m <= 64) {  // insertion sort
    for (int i = 1; i < m; i++) {
      const int v = a[i];
      int j = i - 1;
      while (j >= 0 && a[j] > v) {
        a[j + 1] = a[j];
        j--;
      }
      a[j + 1] = v;
    }
  } else {  // heap sort (edge cells collecting escaped particles)
    for (int start = m / 2 - 1; start >= 0; start--) {
      int root = start;
      for (;;) {
        int child = 2 * root + 1;
        if (child >= m) break;
        if (child + 1 < m && a[child] < a[child + 1]) child++;
        if (a[root] >= a[child]) break;
        const int t = a[root]; a[root] = a[child]; a[child] = t;
        root = child;
      }
    }
    for (int end = m - 1; end > 0; end--) {
      const int t0 = a[0]; a[0] = a[end]; a[end] = t0;
      int root = 0;
      for (;;) {
        int child = 2 * root + 1;
        if (child >= end) break;
        if (child + 1 < end && a[child] < a[child + 1]) child++;
        if (a[root] >= a[child]) break;
        const int t = a[root]; a[root] = a[child]; a[child] = t;
        root = child;
      }
    }
  }
}
// Same ordering pass driven from the particles: the particle that drew rank 0 of a cell sorts that cell's slice
// (one thread per particle instead of one per cell: the grid covers the whole tank, most of its cells are empty)
__global__ void k_bin_sort_cells_by_particle(const int *n_ptr, int n_fixed, const unsigned int *cell_start, const int *cell_of_particle,
                                             const int *rank_in_cell, int *sorted_src) {
  const int n = n_ptr ? *n_ptr : n_fixed;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || rank_in_cell[i] != 0) return;
  const int c = cell_of_particle[i];
  const int s = (int)cell_start[c], m = (int)cell_start[c + 1] - s;
  if (m < 2) return;
  int *a = sorted_src + s;
  if (m <= 64) {
    for (int p = 1; p < m; p++) {
      const int v = a[p];
      int j = p - 1;
      while (j >= 0 && a[j] > v) {
        a[j + 1] = a[j];
        j--;
      }
      a[j + 1] = v;
    }
  } else {  // heap sort (edge cells collecting escaped particles)
    for (int start = m / 2 - 1; start >= 0; start--) {
      int root = start;
      for (;;) {
        int child = 2 * root + 1;
        if (child >= m) break;
        if (child + 1 < m && a[child] < a[child + 1]) child++;
        if (a[root] >= a[child]) break;
        const int t = a[root]; a[root] = a[child]; a[child] = t;
        root = child;
      }
    }
    for (int end = m - 1; end > 0; end--) {
      const int t0 = a[0]; a[0] = a[end]; a[end] = t0;
      int root = 0;
      for (;;) {
        int child = 2 * root + 1;
        if (child >= end) break;
        if (child + 1 < end && a[child] < a[child + 1]) child++;
        if (a[root] >= a[child]) break;
        const int t = a[root]; a[root] = a[child]; a[child] = t;
        root = child;
      }
    }
  }
}
// gather every persistent per-particle array into cell order (replaces FluidModel::
// performNeighborhoodSearchSort + SimulationDataDiffDFSPH::performNeighborhoodSearchSort,
// FluidModel.cpp:357-386, SimulationDataDiffDFSPH.cpp:150-171 — done every step here)
__global__ void k_permute_fluid(const StepState *st, const int *sorted_src, const double4 *pos_in, const double4 *vel_in,
                                const double *kappa_in, const double *kappav_in, const int *id_in, const int *state_in,
                                double4 *pos_out, double4 *vel_out, double *kappa_out, double *kappav_out, int *id_out,
                                int *state_out, const int *src_base = nullptr) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= st->nf) return;
  const int s = sorted_src[i] + (src_base ? *src_base : 0);  // slab mode: the source range starts at own_begin
  pos_out[i] = pos_in[s];
  vel_out[i] = vel_in[s];
  kappa_out[i] = kappa_in[s];
  kappav_out[i] = kappav_in[s];
  id_out[i] = id_in[s];
  state_out[i] = state_in[s];
}
// xyz AoS (what crosses the C ABI) -> double4 records
__global__ void k_xyz_to_rec(const double *xyz, double4 *out, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = make_double4(xyz[3 * (size_t)i], xyz[3 * (size_t)i + 1], xyz[3 * (size_t)i + 2], 0.0);
}
// one-time physical sort of the static boundary particles
__global__ void k_permute_boundary(int n, const int *sorted_src, const double4 *pos_in, const double4 *x0_in, const int *body_in,
                                   const int *orig_in, double4 *pos_out, double4 *x0_out, int *body_out, int *orig_out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int s = sorted_src[i];
  pos_out[i] = pos_in[s];
  x0_out[i] = x0_in[s];
  body_out[i] = body_in[s];
  orig_out[i] = orig_in[s];
}

// "A point of this set may be within reach": one byte per cell, set for every cell of the (2 reach + 1)^3 block around
// the cell of each point - exactly the cells whose stencil walk (for_each_in_range) would visit that point's cell.
// The list build reads the byte of a fluid particle's own cell and skips the whole 25-row walk over the static
// boundary set when it is 0 (true for most of the fluid).  Marked once for the static set (those particles never move)
// and with every rebuild of the dynamic set's cell table (k_mark_near_warp: few particles, so one warp per particle).
__global__ void k_mark_near_warp(const __grid_constant__ Params P, const double4 *pos, int n, unsigned char *near_flag) {
  const int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (t >= n) return;
  const double4 p = pos[t];
  int cx, cy, cz;
  cell_of(P.grid, p.x, p.y, p.z, cx, cy, cz);
  const int R = P.grid.reach, W = 2 * R + 1;
  for (int k = lane; k < W * W * W; k += 32) {
    const int x = cx - R + k % W, y = cy - R + (k / W) % W, z = cz - R + k / (W * W);
    if (x >= 0 && x < P.grid.nx && y >= 0 && y < P.grid.ny && z >= 0 && z < P.grid.nz) near_flag[cell_lin(P.grid, x, y, z)] = 1;
  }
}
__global__ void k_mark_near(const __grid_constant__ Params P, const double4 *pos, int n, unsigned char *near_flag) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const double4 p = pos[t];
  int cx, cy, cz;
  cell_of(P.grid, p.x, p.y, p.z, cx, cy, cz);
  const int R = P.grid.reach;
  for (int z = max(cz - R, 0); z <= min(cz + R, P.grid.nz - 1); z++)
    for (int y = max(cy - R, 0); y <= min(cy + R, P.grid.ny - 1); y++)
      for (int x = max(cx - R, 0); x <= min(cx + R, P.grid.nx - 1); x++) near_flag[cell_lin(P.grid, x, y, z)] = 1;
}

// ---------------------------------------------------------------------------------------------
// neighbour lists
// ---------------------------------------------------------------------------------------------
struct GridView {
  const unsigned int *cell_start;  // ncells + 1
  const int *sorted_src;           // nullptr: points are physically in cell order; else slot -> point (relative)
  const double4 *pos;              // points of this set
  int base;                        // added to the point index stored in the list
};

// Visits every point of `g` within the support radius of (px,py,pz); f(index) with index = base + point.
// -DDFR_NBR_TRIM=1 trims the (2 reach + 1)^3 stencil to the cells the support sphere can reach: a row whose cell slab
// lies further away than the radius in (y, z) is skipped, the others are cut in x to the cells whose near face is still
// inside (squared face distances, 1e-9 relative slack - far more than the rounding of cell_of; points outside the grid
// sit in edge cells, whose far side is open, so the near-face test holds for them too).  84 of 125 cells and 21.6 of
// 25 rows are left for reach 2, order and set of the accepted points are unchanged (the parity tests pass) - and
// k_nbr_build gets SLOWER, 386 -> 434 us at 1 M particles: a warp still runs the longest row of its 32 lanes, so the
// trip counts barely fall while every row pays the face arithmetic.  Off by default (profiles/r2_gather_analysis.md).
#ifndef DFR_NBR_TRIM
#define DFR_NBR_TRIM 0
#endif
template <class F>
__device__ __forceinline__ void for_each_in_range(const Params &P, const GridView &g, double px, double py, double pz, int self, F f) {
  int cx, cy, cz;
  cell_of(P.grid, px, py, pz, cx, cy, cz);
  const int R = P.grid.reach;
  const int xlo = max(cx - R, 0), xhi = min(cx + R, P.grid.nx - 1);
#if DFR_NBR_TRIM
  const double cell = 1.0 / P.grid.inv_cell;
  const double r2m = P.r2 * (1.0 + 1e-9);
  const double ry = py - P.grid.oy, rz = pz - P.grid.oz, rx = px - P.grid.ox;
#endif
  for (int z = max(cz - R, 0); z <= min(cz + R, P.grid.nz - 1); z++) {
#if DFR_NBR_TRIM
    // distance to the slab of cell layer z (0 for my own layer and when I sit beyond the face)
    const double dz = z < cz ? rz - (double)(z + P.grid.z_shift + 1) * cell : (z > cz ? (double)(z + P.grid.z_shift) * cell - rz : 0.0);
    const double dz2 = dz > 0.0 ? dz * dz : 0.0;
    if (dz2 >= r2m) continue;
#endif
    for (int y = max(cy - R, 0); y <= min(cy + R, P.grid.ny - 1); y++) {
      int xl = xlo, xh = xhi;
#if DFR_NBR_TRIM
      const double dy = y < cy ? ry - (double)(y + 1) * cell : (y > cy ? (double)y * cell - ry : 0.0);
      const double d2 = dz2 + (dy > 0.0 ? dy * dy : 0.0);
      if (d2 >= r2m) continue;
      for (xl = cx; xl > xlo; xl--) {  // take cell xl - 1 while its high face is in reach
        const double d = rx - (double)xl * cell;
        if (d > 0.0 && d * d + d2 >= r2m) break;
      }
      for (xh = cx; xh < xhi; xh++) {
        const double d = (double)(xh + 1) * cell - rx;
        if (d > 0.0 && d * d + d2 >= r2m) break;
      }
#endif
      const int s = (int)g.cell_start[cell_lin(P.grid, xl, y, z)];
      const int e = (int)g.cell_start[cell_lin(P.grid, xh, y, z) + 1];
      for (int p = s; p < e; p++) {
        const int j = g.sorted_src ? g.sorted_src[p] : p;
        if (j == self) continue;
        const double4 q = ldg4(g.pos + j);
        if (dist2_exact(px, py, pz, q.x, q.y, q.z) < P.r2) f(g.base + j);
      }
    }
  }
}

// Single scan: fluid->fluid and fluid->boundary neighbour lists in a warp-interleaved ELL layout
// (dfr_types.cuh: "ELL-4"; no count pass / prefix sum is needed).  Rows longer than cap raise error_flags and report
// their length; the host then grows the capacity and rebuilds the lists (dfr_api.cu: ensure_list_capacity).
__global__ void __launch_bounds__(DFR_CTA_THREADS) k_nbr_build(const __grid_constant__ Params P, StepState *st, const double4 *pos, GridView gf,
                                                    GridView gs, GridView gd, int has_static, int has_dyn, int *cnt_f, int *cnt_b,
                                                    int *idx_f, int *idx_b, int cap_f, int cap_b, const unsigned char *near_s,
                                                    const unsigned char *near_d, const VSched S) {
  vsched_prologue(S);
  const int n = st->nf;
  DFR_VB_LOOP(S) {
  const int i = vb_ * 128 + DFR_TID;
  int cf = 0, cb = 0;
  (void)n;
  if (i >= st->own_begin && i < st->own_end) {
    const int lane = i & 31;
    const double4 p = pos[i];
    (void)lane;
    // four indices are collected in registers and leave as one 16-byte store (the ELL-4 group of this lane): a quarter of
    // the store instructions of the scalar version, each a full group instead of a 4-byte piece of it
    int4 pend = make_int4(0, 0, 0, 0);
    int4 *row_f = reinterpret_cast<int4 *>(idx_f) + ((size_t)(i >> 5) * (size_t)(cap_f >> 2)) * 32 + (i & 31);
    for_each_in_range(P, gf, p.x, p.y, p.z, i, [&](int j) {
      const int k = cf & 3;
      if (k == 0) pend.x = j;
      else if (k == 1) pend.y = j;
      else if (k == 2) pend.z = j;
      else {
        pend.w = j;
        if (cf < cap_f) row_f[(size_t)(cf >> 2) * 32] = pend;
      }
      cf++;
    });
    if ((cf & 3) != 0 && cf < cap_f) row_f[(size_t)(cf >> 2) * 32] = pend;  // the last, partly filled group
    int ocx, ocy, ocz;
    cell_of(P.grid, p.x, p.y, p.z, ocx, ocy, ocz);
    const int own_cell = cell_lin(P.grid, ocx, ocy, ocz);
    if (has_static && near_s[own_cell]) for_each_in_range(P, gs, p.x, p.y, p.z, -1, [&](int j) {
      if (cb < cap_b) idx_b[nbr_slot(cap_b, i, cb)] = j;
      cb++;
    });
    if (has_dyn && near_d[own_cell]) for_each_in_range(P, gd, p.x, p.y, p.z, -1, [&](int j) {
      if (cb < cap_b) idx_b[nbr_slot(cap_b, i, cb)] = j;
      cb++;
    });
    if (cf > cap_f) atomicOr(&st->error_flags, 1);
    if (cb > cap_b) atomicOr(&st->error_flags, 2);
    cnt_f[i] = min(cf, cap_f);
    cnt_b[i] = min(cb, cap_b);
  }
  // neighbour statistics of the step (mean neighbour count feeds the roofline's algorithmic bytes)
  const int tot = __reduce_add_sync(DFR_FULL, cf + cb);
  const int mf = __reduce_max_sync(DFR_FULL, cf), mb = __reduce_max_sync(DFR_FULL, cb);
  if ((threadIdx.x & 31) == 0 && tot) {
    atomicAdd((unsigned long long *)&st->total_neighbors, (unsigned long long)tot);
    if ((unsigned int)mf > st->list_used_f) atomicMax(&st->list_used_f, (unsigned int)mf);
    if ((unsigned int)mb > st->list_used_b) atomicMax(&st->list_used_b, (unsigned int)mb);
  }
  }
}
// One stencil row of a point in cell (cx, cy, cz): row rr = (z - (cz - R)) * (2R + 1) + (y - (cy - R)), the order in
// which for_each_in_range walks the rows; [s, e) is the slot range of the row's cells (empty outside the grid).
__device__ __forceinline__ void stencil_row(const Params &P, const GridView &g, int cx, int cy, int cz, int rr, int &s, int &e) {
  const int R = P.grid.reach, W = 2 * R + 1;
  const int z = cz - R + rr / W, y = cy - R + rr % W;
  s = e = 0;
  if (z < 0 || z >= P.grid.nz || y < 0 || y >= P.grid.ny) return;
  const int xlo = max(cx - R, 0), xhi = min(cx + R, P.grid.nx - 1);
  s = (int)g.cell_start[cell_lin(P.grid, xlo, y, z)];
  e = (int)g.cell_start[cell_lin(P.grid, xhi, y, z) + 1];
}
// dynamic boundary particle -> fluid neighbours, CSR rows (one warp later walks one row).  There are few dynamic
// boundary particles (hundreds to ~15 k), so one WARP searches for one particle, a lane per stencil row; rows are
// written in stencil order, i.e. exactly the order a sequential walk produces.
__global__ void __launch_bounds__(128) k_dnbr_count(const __grid_constant__ Params P, const double4 *bpos, int dyn_begin, int n_dyn, GridView gf, unsigned int *cnt_d) {
  const int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (t >= n_dyn) return;
  const double4 p = bpos[dyn_begin + t];
  int cx, cy, cz;
  cell_of(P.grid, p.x, p.y, p.z, cx, cy, cz);
  const int rows = (2 * P.grid.reach + 1) * (2 * P.grid.reach + 1);
  unsigned int c = 0;
  for (int rr = lane; rr < rows; rr += 32) {
    int s, e;
    stencil_row(P, gf, cx, cy, cz, rr, s, e);
    for (int q = s; q < e; q++) {
      const int j = gf.sorted_src ? gf.sorted_src[q] : q;
      const double4 x = ldg4(gf.pos + j);
      if (dist2_exact(p.x, p.y, p.z, x.x, x.y, x.z) < P.r2) c++;
    }
  }
  c = __reduce_add_sync(DFR_FULL, c);
  if (lane == 0) cnt_d[t] = c;
}
__global__ void __launch_bounds__(128) k_dnbr_fill(const __grid_constant__ Params P, StepState *st, const double4 *bpos, int dyn_begin, int n_dyn, GridView gf,
                                                    const unsigned int *off_d, int *idx_d, unsigned int cap_d) {
  const int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (t >= n_dyn) return;
  if (t == 0 && lane == 0) st->list_used_d = off_d[n_dyn];
  if (off_d[t + 1] > cap_d) {
    if (lane == 0) atomicOr(&st->error_flags, 4);
    return;
  }
  const double4 p = bpos[dyn_begin + t];
  int cx, cy, cz;
  cell_of(P.grid, p.x, p.y, p.z, cx, cy, cz);
  const int rows = (2 * P.grid.reach + 1) * (2 * P.grid.reach + 1);
  unsigned int base = off_d[t];
  for (int r0 = 0; r0 < rows; r0 += 32) {
    const int rr = r0 + lane;
    int s = 0, e = 0;
    if (rr < rows) stencil_row(P, gf, cx, cy, cz, rr, s, e);
    unsigned int m = 0;
    for (int q = s; q < e; q++) {
      const int j = gf.sorted_src ? gf.sorted_src[q] : q;
      const double4 x = ldg4(gf.pos + j);
      if (dist2_exact(p.x, p.y, p.z, x.x, x.y, x.z) < P.r2) m++;
    }
    unsigned int incl = m;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned int v = __shfl_up_sync(DFR_FULL, incl, o);
      if (lane >= o) incl += v;
    }
    int *o = idx_d + base + (incl - m);
    for (int q = s; q < e; q++) {
      const int j = gf.sorted_src ? gf.sorted_src[q] : q;
      const double4 x = ldg4(gf.pos + j);
      if (dist2_exact(p.x, p.y, p.z, x.x, x.y, x.z) < P.r2) *o++ = gf.base + j;
    }
    base += __shfl_sync(DFR_FULL, incl, 31);
  }
}

// ---------------------------------------------------------------------------------------------
// boundary volume psi (Simulation::updateBoundaryVolume, Simulation.cpp:831-902;
// BoundaryModel_Akinci2012::computeBoundaryVolume, BoundaryModel_Akinci2012.cpp:257-284)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_boundary_volume(const __grid_constant__ Params P, double4 *bpos, int n_b, int n_static, GridView gs, GridView gd,
                                                          int has_static, int has_dyn, double *vol_out) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= n_b) return;
  const double4 p = bpos[b];
  double delta = P.W_zero;
  const int self_s = (b < n_static) ? b : -1;
  const int self_d = (b >= n_static) ? b - n_static : -1;
  if (has_static)
    for_each_in_range(P, gs, p.x, p.y, p.z, self_s, [&](int j) {
      const double4 q = ldg4(bpos + j);
      const double dx = p.x - q.x, dy = p.y - q.y, dz = p.z - q.z;
      delta += cubic_W(P, dx * dx + dy * dy + dz * dz);
    });
  if (has_dyn)
    for_each_in_range(P, gd, p.x, p.y, p.z, self_d, [&](int j) {
      const double4 q = ldg4(bpos + j);
      const double dx = p.x - q.x, dy = p.y - q.y, dz = p.z - q.z;
      delta += cubic_W(P, dx * dx + dy * dy + dz * dz);
    });
  vol_out[b] = 1.0 / delta;
}
__global__ void k_store_volume(double4 *bpos, const double *vol, int n_b) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < n_b) bpos[b].w = vol[b];
}

// ---------------------------------------------------------------------------------------------
// density + DFSPH factor, fused (TimeStep::computeDensities, TimeStep.cpp:147-200;
// computeDFSPHFactor, TimeStepDiffDFSPH.cpp:883-962)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(DFR_CTA_THREADS) k_density_factor(const __grid_constant__ Params P, const StepState *st, const double4 *pos, const double4 *bpos,
                                                         NbrList lf, NbrList lb, double *density, double *factor, double4 *sgp, double4 *xrho,
                                                         const GhostOut GO, const VSched S) {
  vsched_prologue(S);
  DFR_VB_LOOP(S) {
  const int i = vb_ * 128 + DFR_TID;
  if (i < st->own_begin || i >= st->own_end) continue;
  const double4 pi = pos[i];
  double dens = P.volume * P.W_zero;
  double S = 0.0;
  d3 G = mk3(0, 0, 0);  // sum_j V_j gradW_ij
  for_neighbors4<DFR_DF_U>(
      lf, i, i, [&](int j) { return ldg4(pos + j); },
      [&](const double4 &pj, int) {
        const d3 r = mk3(pi.x - pj.x, pi.y - pj.y, pi.z - pj.z);
        double c;
        const double wv = cubic_W_and_grad(P, dot(r, r), c);
        dens += P.volume * wv;
        const d3 g = (P.volume * c) * r;
        S += dot(g, g);
        G += g;
      });
  for_neighbors4(
      lb, i, 0, [&](int j) { return ldg4(bpos + j); },
      [&](const double4 &pj, int) {
        const d3 r = mk3(pi.x - pj.x, pi.y - pj.y, pi.z - pj.z);
        double c;
        const double wv = cubic_W_and_grad(P, dot(r, r), c);
        dens += pj.w * wv;
        G += (pj.w * c) * r;
      });
  density[i] = dens * P.density0;
  // (x, rho) record: k_normals gathers position and density of a neighbour in one 256-bit load
  stg4(xrho + i, make_double4(pi.x, pi.y, pi.z, dens * P.density0));
  ghost_store(GO, i, make_double4(pi.x, pi.y, pi.z, dens * P.density0));
  const double denom = S + dot(G, G);
  factor[i] = (denom > DFR_EPS) ? -1.0 / denom : 0.0;
  sgp[i] = make_double4(-G.x, -G.y, -G.z, 0.0);
  }
}

// ---------------------------------------------------------------------------------------------
// velocity-divergence pass: computeDensityChange / computeDensityAdv (+ warm-start clamps, the
// stiffness for the next push, and the residual reduction with on-device convergence control)
// TimeStepDiffDFSPH.cpp:1926-2041, 964-999, 1698-1732, 1203-1216, 1906-1915, 711-743, 828-861
// ---------------------------------------------------------------------------------------------
enum { RHO_PLAIN = 0, RHO_WARM = 1, RHO_ITER = 2 };
// Passes fused into the first two k_rho launches of a step (divergence solve with warm start; the gathers are the
// same, so a fused pass costs FP64 work but no extra trip through the neighbour lists):
//   RHO_X_DENSITY  the first launch also computes density, DFSPH factor, sum V gradW and the (x, rho) record
//                  (k_density_factor is then not launched)
//   RHO_X_NORMALS  the second launch gathers (x, rho) instead of x and also computes the surface-tension normals
//                  (k_normals is then not launched); positions and densities do not change during the solve
//   RHO_X_NONPRESSURE  an iteration launch also evaluates the non-pressure accelerations (surface tension + viscosity,
//                  k_nonpressure's pair loop: it gathers (x, rho), v and the normal) into `acc`.  The result is only
//                  valid if that launch turns out to be the LAST active iteration of the divergence solve (its input
//                  velocities are then the final ones); the host checks the iteration count and otherwise runs
//                  k_nonpressure.  k_apply_accel then does v += h a, the CFL maximum and the kappa_v rescale.
enum { RHO_X_NONE = 0, RHO_X_DENSITY = 1, RHO_X_NORMALS = 2, RHO_X_NONPRESSURE = 3 };
struct RhoExtra {
  double *density, *factor;
  double4 *sgp, *xrho, *normal, *acc;
  GhostOut go;  // ghost rows of the extra gathered array (xrho or normal)
  int gate;     // graph stepping: 0 always run, 1 run only if st->fuse_now, 2 run only if !st->fuse_now
};
// pair terms of SurfaceTension_Akinci2013::step (SurfaceTension_Akinci2013.cpp:60-151) and Viscosity_Standard::step
// (Viscosity_Standard.cpp:233-334) for one fluid neighbour; r = x_i - x_j, c the cubic gradient coefficient
__device__ __forceinline__ void nonpressure_pair(const Params &P, bool st_on, bool visc_on, const d3 &r, double r2, double c, double rhoi,
                                                 double rhoj, const d3 &ni, const double4 &nj, double vdotr, double h2s, d3 &a) {
  if (st_on) {
    const double K_ij = 2.0 * P.density0 / (rhoi + rhoj);
    d3 accel = mk3(0, 0, 0);
    if (r2 > 1.0e-9) accel -= (P.surface_tension * P.mass * cohesion_W(P, r2) * rsqrt(r2)) * r;
    accel -= P.surface_tension * mk3(ni.x - nj.x, ni.y - nj.y, ni.z - nj.z);
    a += K_ij * accel;
  }
  if (visc_on) a += (10.0 * P.viscosity * (P.mass / rhoj) * vdotr / (r2 + 0.01 * h2s) * c) * r;
}

template <bool PRESSURE, int MODE, int EXTRA = RHO_X_NONE>
__global__ void __launch_bounds__(DFR_CTA_THREADS, DFR_RESIDENT(EXTRA == RHO_X_NONE ? DFR_RHO_BLOCKS : (EXTRA == RHO_X_NONPRESSURE ? DFR_RHONP_BLOCKS : DFR_RHOX_BLOCKS))) k_rho(const __grid_constant__ Params P, StepState *st, const double4 *pos, const double4 *vel, const double4 *bpos,
                                              const double4 *bvel, NbrList lf, NbrList lb, const double *density, const double *factor,
                                              const int *state, double *kappa, double *dadv, double4 *xk, double *partials, const GhostOut GO,
                                              const RhoExtra X, const VSched S) {
  vsched_prologue(S);
  if (MODE == RHO_ITER) {
    if (!(PRESSURE ? st->prs_active : st->div_active)) return;
    // graph stepping launches the plain and the fused variant of every divergence iteration; one of them runs
    if (X.gate == 1 && !st->fuse_now) return;
    if (X.gate == 2 && st->fuse_now) return;
  }
  const int nf = st->nf;
  DFR_VB_LOOP(S) {
  if (vb_ * 128 >= nf) continue;  // grids are sized for the emitter capacity; only live blocks take a ticket
  const int i = vb_ * 128 + DFR_TID;
  const double h = PRESSURE ? st->h : st->h_step;
  double err = 0.0;
  if (i >= st->own_begin && i < st->own_end) {
    const double4 pi = pos[i];
    const double4 vi = vel[i];
    double delta = 0.0;
    // fused extras
    double dens = P.volume * P.W_zero, Ssum = 0.0;
    d3 G = mk3(0, 0, 0);  // sum_j V_j gradW_ij (RHO_X_DENSITY) / normal sum (RHO_X_NORMALS)
    const int nF = lf.cnt[i];
    // RHO_X_NONPRESSURE state (pos == xrho: pi.w is the particle's density)
    const bool st_on = (EXTRA == RHO_X_NONPRESSURE) && P.st_method == 2;
    const bool visc_on = (EXTRA == RHO_X_NONPRESSURE) && P.visc_method == 1;
    const double h2s = P.support_radius * P.support_radius;
    d3 ni = mk3(0, 0, 0), acc = mk3(0, 0, 0);
    if (EXTRA == RHO_X_NONPRESSURE) {
      if (st_on) {
        const double4 n4 = X.normal[i];
        ni = mk3(n4.x, n4.y, n4.z);
      }
      for_neighbors4<DFR_NP_U>(
          lf, i, i,
          [&](int j) {
            Rec3 q;
            q.a = ldg4(pos + j);
            q.b = ldg4(vel + j);
            if (st_on) q.c = ldg4(X.normal + j);
            return q;
          },
          [&](const Rec3 &q, int) {
            const d3 r = mk3(pi.x - q.a.x, pi.y - q.a.y, pi.z - q.a.z);
            const double r2 = dot(r, r);
            const double c = cubic_grad_coeff(P, r2);
            const double vdotr = (vi.x - q.b.x) * r.x + (vi.y - q.b.y) * r.y + (vi.z - q.b.z) * r.z;
            delta += (P.volume * c) * vdotr;
            nonpressure_pair(P, st_on, visc_on, r, r2, c, pi.w, q.a.w, ni, q.c, vdotr, h2s, acc);
          });
    } else
    for_neighbors4<DFR_RHO_U>(
        lf, i, i, [&](int j) { return Rec2{ldg4(pos + j), ldg4(vel + j)}; },
        [&](const Rec2 &q, int) {
          const d3 r = mk3(pi.x - q.a.x, pi.y - q.a.y, pi.z - q.a.z);
          double c;
          if (EXTRA == RHO_X_DENSITY) {
            const double wv = cubic_W_and_grad(P, dot(r, r), c);
            dens += P.volume * wv;
            const d3 g = (P.volume * c) * r;
            Ssum += dot(g, g);
            G += g;
          } else {
            c = cubic_grad_coeff(P, dot(r, r));
            if (EXTRA == RHO_X_NORMALS) G += (P.mass / q.a.w * c) * r;  // pos == xrho here: w is the neighbour's density
          }
          delta += (P.volume * c) * ((vi.x - q.b.x) * r.x + (vi.y - q.b.y) * r.y + (vi.z - q.b.z) * r.z);
        });
    const int nB = lb.cnt[i];
    for_neighbors4<DFR_RHO_U>(
        lb, i, 0, [&](int j) { return Rec2{ldg4(bpos + j), ldg4(bvel + j)}; },
        [&](const Rec2 &q, int) {
          const d3 r = mk3(pi.x - q.a.x, pi.y - q.a.y, pi.z - q.a.z);
          double c;
          if (EXTRA == RHO_X_DENSITY) {
            const double wv = cubic_W_and_grad(P, dot(r, r), c);
            dens += q.a.w * wv;
            G += (q.a.w * c) * r;
          } else
            c = cubic_grad_coeff(P, dot(r, r));
          const double vdotr = (vi.x - q.b.x) * r.x + (vi.y - q.b.y) * r.y + (vi.z - q.b.z) * r.z;
          delta += (q.a.w * c) * vdotr;
          if (EXTRA == RHO_X_NONPRESSURE) {  // boundary adhesion / boundary viscosity (zero coefficients in every shipped scene)
            const double r2 = dot(r, r);
            if (st_on && P.surface_tension_b != 0.0 && r2 > 1.0e-9)
              acc -= (P.surface_tension_b * P.density0 * q.a.w * adhesion_W(P, r2) * rsqrt(r2)) * r;
            if (visc_on && P.viscosity_b != 0.0)
              acc += (10.0 * P.viscosity_b * (P.density0 * q.a.w / pi.w) * vdotr / (r2 + 0.01 * h2s) * c) * r;
          }
        });
    if (EXTRA == RHO_X_NONPRESSURE) X.acc[i] = make_double4(P.gx + acc.x, P.gy + acc.y, P.gz + acc.z, 0.0);
    double alpha_i = 0.0;
    if (EXTRA == RHO_X_DENSITY) {  // k_density_factor's epilogue
      X.density[i] = dens * P.density0;
      stg4(X.xrho + i, make_double4(pi.x, pi.y, pi.z, dens * P.density0));
      ghost_store(X.go, i, make_double4(pi.x, pi.y, pi.z, dens * P.density0));
      const double denom = Ssum + dot(G, G);
      alpha_i = (denom > DFR_EPS) ? -1.0 / denom : 0.0;
      X.factor[i] = alpha_i;
      X.sgp[i] = make_double4(-G.x, -G.y, -G.z, 0.0);
    } else if (MODE != RHO_WARM)
      alpha_i = factor[i];
    if (EXTRA == RHO_X_NORMALS) {  // k_normals' epilogue
      const double4 nrec = make_double4(P.support_radius * G.x, P.support_radius * G.y, P.support_radius * G.z, pi.w);
      stg4(X.normal + i, nrec);
      ghost_store(X.go, i, nrec);
    }
    double rho;
    if (PRESSURE) {
      rho = fmax(density[i] / P.density0 + h * delta, 1.0);  // (the pressure solve never carries an extra)
      err = P.density0 * rho - P.density0;
    } else {
      rho = fmax(delta, 0.0);
      if (nF + nB < 20) rho = 0.0;  // particle deficiency (TimeStepDiffDFSPH.cpp:2027-2040)
      err = P.density0 * rho;
    }
    dadv[i] = rho;
    if (MODE == RHO_WARM) {
      double kap;
      if (PRESSURE)
        kap = (rho > 1.0) ? 0.5 * fmax(kappa[i], -0.00025) / (h * h) : 0.0;
      else
        kap = (rho > 0.0) ? 0.5 * fmax(kappa[i], -0.5) / h : 0.0;
      if (state[i] != 0) kap = 0.0;  // the reference zeroes these in its second loop (:1008-1012 / :1737-1741)
      kappa[i] = kap;
      stg4(xk + i, make_double4(pi.x, pi.y, pi.z, kap));
      ghost_store(GO, i, make_double4(pi.x, pi.y, pi.z, kap));
    } else {
      const double b = PRESSURE ? rho - 1.0 : rho;
      const double ks = PRESSURE ? b * alpha_i / (h * h) : b * alpha_i / h;
      // (x, k) record: the push and the boundary-side kernel gather position and stiffness together
      stg4(xk + i, make_double4(pi.x, pi.y, pi.z, ks));
      ghost_store(GO, i, make_double4(pi.x, pi.y, pi.z, ks));
    }
  }
  if (MODE == RHO_ITER) {
    // deterministic residual: one partial per warp, summed in a fixed order by k_residual_finish (no shared memory,
    // barrier, fence or ticket here: they cost every iteration pass ~20 us against the plain passes)
    double s = err;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(DFR_FULL, s, o);
    if ((threadIdx.x & 31) == 0) partials[(size_t)vb_ * 4 + (DFR_TID >> 5)] = s;
  }
  }
}

// Closes a Jacobi iteration: sums the per-warp residual partials of k_rho<..., RHO_ITER> in a fixed order and applies
// the stopping rules of pressureSolve / divergenceSolve (TimeStepDiffDFSPH.cpp:711-743, :828-861).  RES_BLOCKS blocks
// each sum one contiguous slice (fixed assignment, fixed tree), the block that finishes last adds the slice sums in
// slice order and decides - the result does not depend on which block that is.
// Slab-decomposed contexts only store the local sum: the rule needs the sum over all slabs (k_solver_decide).
#define RES_THREADS 256
#define RES_BLOCKS 16
template <bool PRESSURE>
__device__ __forceinline__ void solver_decide(const Params &P, StepState *st, double avg, unsigned long long cond, int policy);
// cond / policy (graph stepping, else 0): `cond` is the handle of the WHILE node this kernel's iteration body hangs in -
// it is cleared when the solve closes; policy 1 = this is the divergence solve with the fused non-pressure pass enabled:
// decide whether the next iteration carries it (the iteration count of the previous step predicts the last iteration;
// once the prediction is exceeded every further iteration probably is the last) and whether the pass that just ran was
// the last one (np_done).
template <bool PRESSURE>
__global__ void __launch_bounds__(RES_THREADS) k_residual_finish(const __grid_constant__ Params P, StepState *st, const double *partials,
                                                                 double *slice_sums, unsigned long long cond, int policy) {
  if (!(PRESSURE ? st->prs_active : st->div_active)) {
    if (cond && blockIdx.x == 0 && threadIdx.x == 0) cudaGraphSetConditional((cudaGraphConditionalHandle)cond, 0u);
    return;
  }
  const int nf = st->nf;
  const int np = ((nf + 127) / 128) * 4;
  const int per = (np + RES_BLOCKS - 1) / RES_BLOCKS;
  const int lo = blockIdx.x * per, hi = min(lo + per, np);
  // fixed assignment of partials to threads; four independent chains per thread so that the reads overlap
  double a4[4] = {0.0, 0.0, 0.0, 0.0};
  int b = lo + threadIdx.x;
  for (; b + 3 * RES_THREADS < hi; b += 4 * RES_THREADS) {
#pragma unroll
    for (int u = 0; u < 4; u++) a4[u] += partials[b + u * RES_THREADS];
  }
  for (; b < hi; b += RES_THREADS) a4[0] += partials[b];
  double v = (a4[0] + a4[1]) + (a4[2] + a4[3]);
  __shared__ double wsum[RES_THREADS / 32];
  __shared__ bool is_last;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(DFR_FULL, v, o);
  if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < RES_THREADS / 32; w++) t += wsum[w];
    slice_sums[blockIdx.x] = t;
    __threadfence();
    is_last = (atomicAdd(&st->ticket, 1u) == RES_BLOCKS - 1);
  }
  __syncthreads();
  if (!is_last || threadIdx.x != 0) return;
  __threadfence();
  st->ticket = 0;
  double total = 0.0;
#pragma unroll
  for (int k = 0; k < RES_BLOCKS; k++) total += __ldcg(slice_sums + k);
  if (P.slab) {
    st->res_sum = total;
    return;
  }
  solver_decide<PRESSURE>(P, st, total / (double)nf, cond, policy);
}
// The stopping rules of pressureSolve / divergenceSolve (TimeStepDiffDFSPH.cpp:711-743, :828-861) on the mean residual of
// the iteration that just ran; one thread.  cond / policy: see k_residual_finish.
template <bool PRESSURE>
__device__ __forceinline__ void solver_decide(const Params &P, StepState *st, double avg, unsigned long long cond, int policy) {
  st->last_residual = avg;
  if (PRESSURE) {
    const double eta = P.max_error * 0.01 * P.density0;
    const int it = st->prs_iters + 1;
    st->prs_iters = it;
    const bool chk = (avg <= eta);
    const bool go_on = (!chk || it < P.min_iter) && it < P.max_iter;
    if (!go_on) st->prs_active = 0;
    if (cond) cudaGraphSetConditional((cudaGraphConditionalHandle)cond, go_on ? 1u : 0u);
  } else {
    const double eta = (1.0 / st->h_step) * P.max_error_v * 0.01 * P.density0;
    const int it = st->div_iters + 1;
    st->div_iters = it;
    const bool chk = (avg <= eta);
    const bool go_on = (!chk || it < 1) && it < P.max_iter_v;
    if (!go_on) st->div_active = 0;
    if (policy == 1) {
      // the fused pass rides on every iteration from min(iterations of the last two steps) on: a pass that turns out not
      // to be the last costs ~0.6 of a plain pass, a missing one the whole stand-alone k_nonpressure (~2.6 plain passes)
      if (go_on)
        st->fuse_now = (it + 1 >= min(st->spec_div, st->spec_div_prev)) ? 1 : 0;
      else {
        st->np_done = st->fuse_now;  // the pass of this (last) iteration carried the non-pressure accelerations
        st->div_streak = (it == st->spec_div) ? min(st->div_streak + 1, 1000) : 0;
        st->spec_div_prev = st->spec_div;
        st->spec_div = max(it, 1);
      }
    }
    if (cond) cudaGraphSetConditional((cudaGraphConditionalHandle)cond, go_on ? 1u : 0u);
  }
}
// graph stepping: opens the IF node that holds a whole step unless the trajectory has finished (dfr_run_trajectory
// enqueues steps in batches and looks at the state once per batch)
__global__ void k_step_gate(const StepState *st, unsigned long long cond, int respect_finished) {
  cudaGraphSetConditional((cudaGraphConditionalHandle)cond, (respect_finished && st->finished) ? 0u : 1u);
}

// ---------------------------------------------------------------------------------------------
// velocity push (warm starts and Jacobi iterations): TimeStepDiffDFSPH.cpp:1005-1057, 1102-1197,
// 1734-1786, 1826-1902.  Boundary reaction forces are gathered from the boundary side
// (k_boundary_side) instead of being scattered from here.
// ---------------------------------------------------------------------------------------------
template <bool PRESSURE, bool ITER>
__global__ void __launch_bounds__(DFR_CTA_THREADS, DFR_RESIDENT(DFR_PUSH_BLOCKS)) k_push(const __grid_constant__ Params P, const StepState *st, const double4 *xk, double4 *vel, const double4 *bpos,
                                               NbrList lf, NbrList lb, const int *state, double *kappa, int accumulate_kappa,
                                               const GhostOut GO, const VSched S) {
  vsched_prologue(S);
  if (ITER) {
    if (!(PRESSURE ? st->prs_active : st->div_active)) return;
  }
  DFR_VB_LOOP(S) {
  const int i = vb_ * 128 + DFR_TID;
  if (i < st->own_begin || i >= st->own_end) continue;
  if (state[i] != 0) continue;
  const double h = PRESSURE ? st->h : st->h_step;
  const double4 pi = xk[i];
  const double ki = pi.w;
  if (ITER && accumulate_kappa) kappa[i] += ki;
  double4 v = vel[i];
  d3 dv = mk3(0, 0, 0);
  for_neighbors4<DFR_PUSH_U>(
      lf, i, i, [&](int j) { return ldg4(xk + j); },
      [&](const double4 &pj, int) {
        const double kSum = ki + pj.w;
        if (fabs(kSum) > DFR_EPS) {
          const d3 r = mk3(pi.x - pj.x, pi.y - pj.y, pi.z - pj.z);
          const double c = cubic_grad_coeff(P, dot(r, r));
          // vel -= h * kSum * (-V gradW)
          dv += (h * kSum * P.volume * c) * r;
        }
      });
  if (fabs(ki) > DFR_EPS)
    for_neighbors4(
        lb, i, 0, [&](int j) { return ldg4(bpos + j); },
        [&](const double4 &pj, int) {
          const d3 r = mk3(pi.x - pj.x, pi.y - pj.y, pi.z - pj.z);
          const double c = cubic_grad_coeff(P, dot(r, r));
          dv += (h * ki * pj.w * c) * r;
        });
  v.x += dv.x;
  v.y += dv.y;
  v.z += dv.z;
  vel[i] = v;
  ghost_store(GO, i, v);
  }
}

// ---------------------------------------------------------------------------------------------
// boundary side: reaction force/torque on dynamic bodies + the per-pair force/torque Jacobians
// (BoundaryModel::addForce, BoundaryModel.h:46-58; computeRigidBodyGradient / computeGradient,
// TimeStepDiffDFSPH.cpp:1226-1694).  One warp per dynamic boundary particle walks its fluid
// neighbours; 24 pair sums are warp-reduced, lane 0 expands them into the eight Jacobian blocks,
// per-warp rows are combined in a fixed order (deterministic FP64 sums).
// ---------------------------------------------------------------------------------------------
#define BS_WARPS 4
#ifndef BS_PART_PER_BLOCK
#define BS_PART_PER_BLOCK 4  // boundary particles per block (1 per warp): few dynamic particles, so many small blocks
#endif

template <int MODE /*0 pressure, 1 divergence*/, bool GRAD>
__global__ void __launch_bounds__(BS_WARPS * 32) k_boundary_side(const __grid_constant__ Params P, const StepState *st, const BodyDev *bodies, const int *blk_body,
                                                                  const int *blk_first, const double4 *xk, const double4 *vel,
                                                                  const double4 *bpos, const double4 *bvel, const double4 *bx0,
                                                                  int dyn_begin, const unsigned int *off_d, const int *idx_d,
                                                                  const double *dadv, const double *factor,
                                                                  const double4 *sgp, const int *state, int iter_kernel, double *acc_rows) {
  if (iter_kernel) {
    if (!(MODE == 0 ? st->prs_active : st->div_active)) return;
  }
  __shared__ double rows[BS_WARPS][ACC_N];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int k = lane; k < ACC_N; k += 32) rows[wid][k] = 0.0;
  __syncwarp();
  const int body_id = blk_body[blockIdx.x];
  const BodyDev &B = bodies[body_id];
  const bool do_grad = GRAD && !B.animated;
  const double dt = (MODE == 0) ? st->h : st->h_step;  // re-read from the TimeManager at :1230/:1292
  const int first = blk_first[blockIdx.x];
  const int last = min(first + BS_PART_PER_BLOCK, B.p_begin + B.p_count);
  for (int bj = first + wid; bj < last; bj += BS_WARPS) {
    const double4 pj = bpos[bj];
    const double Vj = pj.w;
    const int t = bj - dyn_begin;
    const int s = (int)off_d[t], e = (int)off_d[t + 1];
    double sA[9], sB[9];
    d3 sFhat = mk3(0, 0, 0), sF = mk3(0, 0, 0);
#pragma unroll
    for (int k = 0; k < 9; k++) sA[k] = sB[k] = 0.0;
    d3 vj = mk3(0, 0, 0);
    if (do_grad) {
      const double4 vv = bvel[bj];
      vj = mk3(vv.x, vv.y, vv.z);
    }
    for (int p = s + lane; p < e; p += 32) {
      const int i = idx_d[p];
      if (i < st->own_begin || i >= st->own_end) continue;  // ghosts are summed by the slab that owns them
      const double4 pi = ldg4(xk + i);
      const d3 r = mk3(pi.x - pj.x, pi.y - pj.y, pi.z - pj.z);
      // real reaction force of the push that follows: F = -m dv / dt = m k_i g, g = -V_j gradW
      const double ki_real = pi.w;
      if (!do_grad) {
        if (__ldg(state + i) == 0 && fabs(ki_real) > DFR_EPS) {
          const double c = cubic_grad_coeff(P, dot(r, r));
          sF += (-P.mass * ki_real * Vj * c) * r;
        }
        continue;
      }
      d3 gW;
      m33 H;
      cubic_grad_gradgrad(P, r, gW, H);
      const d3 g = (-Vj) * gW;  // grad_p_j
      if (__ldg(state + i) == 0 && fabs(ki_real) > DFR_EPS) sF += (P.mass * ki_real) * g;
      // ---- computeRigidBodyGradient (:1260-1271) ----
      const double rho = __ldg(dadv + i);
      const double alpha = __ldg(factor + i);
      const double b_i = (MODE == 0) ? rho - 1.0 : rho;
      const double invH = 1.0 / dt, invH2 = 1.0 / dt / dt;
      const unsigned int coeff_trunc = (unsigned int)((MODE == 0) ? invH2 : invH);  // :1266, truncation reproduced
      const double ki = b_i * alpha * (double)coeff_trunc;
      const d3 Fhat = P.mass * ki * g;  // force = -m * (-dt * ki * g) * invH
      sFhat += Fhat;
      // ---- computeGradient (:1281-1482) ----
      const double4 vi4 = ldg4(vel + i);
      const d3 dvij = mk3(vi4.x - vj.x, vi4.y - vj.y, vi4.z - vj.z);
      const bool gate = (MODE == 0) ? (rho > 1.0) : (rho > 0.0);
      const d3 Hdv = H * dvij;
      d3 grad_b_x = mk3(0, 0, 0), grad_b_v = mk3(0, 0, 0), grad_b_vi = mk3(0, 0, 0);
      const double4 sg = ldg4(sgp + i);
      const d3 sumg = mk3(sg.x, sg.y, sg.z);
      if (gate) {
        if (MODE == 0) {
          grad_b_x = (1.0 / P.density0) * g - (dt * Vj) * Hdv;  // :1539-1546 (extra 1/density0 as coded)
          grad_b_v = (-dt * Vj) * gW;                           // :1600
          grad_b_vi = (-dt) * sumg;                             // :1632-1663 == dt * sum_j V_j gradW = -dt * sum_grad_p_k
        } else {
          grad_b_x = (-Vj) * Hdv;  // :1573
          grad_b_v = (-Vj) * gW;   // :1624
          grad_b_vi = -sumg;       // :1665-1694
        }
      }
      d3 grad_alpha_x = mk3(0, 0, 0);
      if (!(alpha >= 0.0)) grad_alpha_x = (2.0 * alpha * alpha * Vj) * (H * sumg);  // :1485-1516
      const double coeff = (MODE == 0) ? invH2 : invH;
      const d3 grad_k_x = coeff * (alpha * grad_b_x + b_i * grad_alpha_x);
      const d3 grad_k_v = (alpha * coeff) * grad_b_v;
      const d3 grad_k_vi = (alpha * coeff) * grad_b_vi;
      // grad_velChange_to_xj = -(g grad_k_x^T + ki * Vj * H)
      m33 Ax = outer(g, grad_k_x);
#pragma unroll
      for (int k = 0; k < 9; k++) Ax.a[k] = -(Ax.a[k] + ki * Vj * H.a[k]);
      m33 Av = outer(g, grad_k_v);
      m33 Avi = outer(g, grad_k_vi);
#pragma unroll
      for (int k = 0; k < 9; k++) {
        Av.a[k] = -Av.a[k];
        Avi.a[k] = -Avi.a[k];
      }
      // dF/dx_j = -m Ax ; dF/dv_j = (I - dt Avi)^-1 (-m) Av   (:1396-1397)
      m33 M = m33::identity();
#pragma unroll
      for (int k = 0; k < 9; k++) M.a[k] -= dt * Avi.a[k];
      const m33 Minv = inverse(M);
      const m33 dFdv = Minv * ((-P.mass) * Av);
#pragma unroll
      for (int k = 0; k < 9; k++) {
        sA[k] += -P.mass * Ax.a[k];
        sB[k] += dFdv.a[k];
      }
    }
    // warp reduction of the 24 pair sums (butterfly: every lane ends with the totals)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      sF.x += __shfl_xor_sync(DFR_FULL, sF.x, o);
      sF.y += __shfl_xor_sync(DFR_FULL, sF.y, o);
      sF.z += __shfl_xor_sync(DFR_FULL, sF.z, o);
    }
    if (do_grad) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int k = 0; k < 9; k++) {
          sA[k] += __shfl_xor_sync(DFR_FULL, sA[k], o);
          sB[k] += __shfl_xor_sync(DFR_FULL, sB[k], o);
        }
        sFhat.x += __shfl_xor_sync(DFR_FULL, sFhat.x, o);
        sFhat.y += __shfl_xor_sync(DFR_FULL, sFhat.y, o);
        sFhat.z += __shfl_xor_sync(DFR_FULL, sFhat.z, o);
      }
    }
    if (lane == 0) {
      double *row = rows[wid];
      const d3 xj = mk3(pj.x, pj.y, pj.z);
      const d3 rj = xj - B.pos;
      row[ACC_F + 0] += sF.x;
      row[ACC_F + 1] += sF.y;
      row[ACC_F + 2] += sF.z;
      const d3 tq = cross(rj, sF);  // torque += (x_j - x_rb) x f  (BoundaryModel.h:56)
      row[ACC_T + 0] += tq.x;
      row[ACC_T + 1] += tq.y;
      row[ACC_T + 2] += tq.z;
      if (do_grad) {
        m33 A, Bm;
#pragma unroll
        for (int k = 0; k < 9; k++) {
          A.a[k] = sA[k];
          Bm.a[k] = sB[k];
        }
#pragma unroll
        for (int k = 0; k < 9; k++) {
          row[ACC_FX + k] += A.a[k];
          row[ACC_FV + k] += Bm.a[k];
        }
        if (P.optimize_rotation) {  // :1425-1478
          const double4 x04 = bx0[bj];
          const d3 r0 = mk3(x04.x, x04.y, x04.z);
          const m34 Q = grad_Rqp_to_q(B.q, r0, -1.0);
          const m33 Sr = skew(rj);
          const m34 dFdq = A * Q + (Bm * skew(B.omega)) * Q;
          const m33 dFdw = Bm * transpose(Sr);
          const m34 dTdq = Sr * dFdq + transpose(skew(sFhat)) * Q;
          const m33 dTdw = Sr * dFdw;
          const m33 dTdv = Sr * Bm;
          const m33 dTdx = Sr * A;
#pragma unroll
          for (int k = 0; k < 12; k++) {
            row[ACC_FQ + k] += dFdq.a[k];
            row[ACC_TQ + k] += dTdq.a[k];
          }
#pragma unroll
          for (int k = 0; k < 9; k++) {
            row[ACC_FW + k] += dFdw.a[k];
            row[ACC_TW + k] += dTdw.a[k];
            row[ACC_TV + k] += dTdv.a[k];
            row[ACC_TX + k] += dTdx.a[k];
          }
        }
      }
    }
    __syncwarp();
  }
  __syncthreads();
  // combine the warps' rows in a fixed order and add them to this block's global row
  for (int k = threadIdx.x; k < ACC_N; k += BS_WARPS * 32) {
    double s = 0.0;
#pragma unroll
    for (int w2 = 0; w2 < BS_WARPS; w2++) s += rows[w2][k];
    acc_rows[(size_t)blockIdx.x * ACC_N + k] += s;
  }
}

// Reaction of the boundary-viscosity term on dynamic bodies (Viscosity_Standard::step, Viscosity_Standard.cpp:273-318):
// a fluid particle i next to boundary particle j gets the acceleration a_ij (k_nonpressure / the fused pass), the body
// gets -m_i a_ij at x_j (BoundaryModel::addForce) and, with BACKWARD defined as the reference builds it, the Jacobian
// -m_i (d a_ij / d v_j + dt d a_ij / d x_j) is added to the particle's dF/dv array (only that one).  Gathered from the
// boundary side like the pressure forces: same blocks, same accumulator rows as k_boundary_side.  Runs after the
// divergence solve (the velocities the non-pressure pass sees) with dt = the step's old time step size.
__global__ void __launch_bounds__(BS_WARPS * 32) k_boundary_viscosity(const __grid_constant__ Params P, const StepState *st, const BodyDev *bodies,
                                                                       const int *blk_body, const int *blk_first, const double4 *xrho,
                                                                       const double4 *vel, const double4 *bpos, const double4 *bvel, int dyn_begin,
                                                                       const unsigned int *off_d, const int *idx_d, double *acc_rows) {
  __shared__ double rows[BS_WARPS][ACC_N];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int k = lane; k < ACC_N; k += 32) rows[wid][k] = 0.0;
  __syncwarp();
  const BodyDev &B = bodies[blk_body[blockIdx.x]];
  const double dt = st->h_step;
  const double h2s = P.support_radius * P.support_radius;
  const int first = blk_first[blockIdx.x];
  const int last = min(first + BS_PART_PER_BLOCK, B.p_begin + B.p_count);
  for (int bj = first + wid; bj < last; bj += BS_WARPS) {
    const double4 pj = bpos[bj];
    const double4 vj4 = bvel[bj];
    const double factor = 10.0 * P.viscosity_b * P.density0 * pj.w;
    const int t = bj - dyn_begin;
    const int s = (int)off_d[t], e = (int)off_d[t + 1];
    d3 sF = mk3(0, 0, 0);
    double sJ[9];
#pragma unroll
    for (int k = 0; k < 9; k++) sJ[k] = 0.0;
    for (int p = s + lane; p < e; p += 32) {
      const int i = idx_d[p];
      if (i < st->own_begin || i >= st->own_end) continue;  // ghosts are summed by the slab that owns them
      const double4 pi = ldg4(xrho + i);  // (x_i, rho_i)
      const double4 vi4 = ldg4(vel + i);
      const d3 r = mk3(pi.x - pj.x, pi.y - pj.y, pi.z - pj.z);
      const d3 dv = mk3(vi4.x - vj4.x, vi4.y - vj4.y, vi4.z - vj4.z);
      d3 gW;
      m33 H;
      cubic_grad_gradgrad(P, r, gW, H);
      const double n = dot(r, r) + 0.01 * h2s;
      const double tmp1 = dot(dv, r) / n;
      const d3 a = (factor / pi.w * tmp1) * gW;
      sF -= P.mass * a;
      // grad1..4 of :283-293, grad_a_to_vj of :303
      const m33 g1 = outer((-1.0 / pi.w * tmp1) * gW, (-1.0) * gW);
      const m33 g2 = outer((1.0 / n) * gW, (-1.0) * dv);
      const m33 g3 = outer((-dot(dv, r) / (n * n) * 2.0) * gW, (-1.0) * r);
      const m33 g_a_v = outer((-factor / pi.w / n) * gW, r);
#pragma unroll
      for (int k = 0; k < 9; k++) {
        const double g_a_x = factor / pi.w * (g1.a[k] + g2.a[k] + g3.a[k] - tmp1 * H.a[k]);
        sJ[k] += -P.mass * (g_a_v.a[k] + dt * g_a_x);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      sF.x += __shfl_xor_sync(DFR_FULL, sF.x, o);
      sF.y += __shfl_xor_sync(DFR_FULL, sF.y, o);
      sF.z += __shfl_xor_sync(DFR_FULL, sF.z, o);
#pragma unroll
      for (int k = 0; k < 9; k++) sJ[k] += __shfl_xor_sync(DFR_FULL, sJ[k], o);
    }
    if (lane == 0) {
      double *row = rows[wid];
      const d3 rj = mk3(pj.x, pj.y, pj.z) - B.pos;
      const d3 tq = cross(rj, sF);
      row[ACC_F + 0] += sF.x;
      row[ACC_F + 1] += sF.y;
      row[ACC_F + 2] += sF.z;
      row[ACC_T + 0] += tq.x;
      row[ACC_T + 1] += tq.y;
      row[ACC_T + 2] += tq.z;
#pragma unroll
      for (int k = 0; k < 9; k++) row[ACC_FV + k] += sJ[k];
    }
    __syncwarp();
  }
  __syncthreads();
  for (int k = threadIdx.x; k < ACC_N; k += BS_WARPS * 32) {
    double s2 = 0.0;
#pragma unroll
    for (int w2 = 0; w2 < BS_WARPS; w2++) s2 += rows[w2][k];
    acc_rows[(size_t)blockIdx.x * ACC_N + k] += s2;
  }
}

// ---------------------------------------------------------------------------------------------
// non-pressure forces: SurfaceTension_Akinci2013::computeNormals / step
// (SurfaceTension_Akinci2013.cpp:25-151), Viscosity_Standard::step (Viscosity_Standard.cpp:233-334),
// clearAccelerations (TimeStep.cpp:66-81), v += h a (TimeStepDiffDFSPH.cpp:589-604) and the CFL
// maximum (Simulation.cpp:542-575), fused.  kappa_v *= h_step of divergenceSolve (:870-880) rides along.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(DFR_CTA_THREADS) k_normals(const __grid_constant__ Params P, const StepState *st, const double4 *xrho, NbrList lf,
                                                  double4 *normal, const GhostOut GO, const VSched S) {
  vsched_prologue(S);
  DFR_VB_LOOP(S) {
  const int i = vb_ * 128 + DFR_TID;
  if (i < st->own_begin || i >= st->own_end) continue;
  const double4 pi = xrho[i];
  d3 n = mk3(0, 0, 0);
  for_neighbors4(
      lf, i, i, [&](int j) { return ldg4(xrho + j); },
      [&](const double4 &pj, int) {
        const d3 r = mk3(pi.x - pj.x, pi.y - pj.y, pi.z - pj.z);
        const double c = cubic_grad_coeff(P, dot(r, r));
        n += (P.mass / pj.w * c) * r;
      });
  // w carries the particle's density so that the force pass gathers (n_j, rho_j) in one record
  stg4(normal + i, make_double4(P.support_radius * n.x, P.support_radius * n.y, P.support_radius * n.z, pi.w));
  ghost_store(GO, i, make_double4(P.support_radius * n.x, P.support_radius * n.y, P.support_radius * n.z, pi.w));
  }
}

__global__ void __launch_bounds__(DFR_CTA_THREADS) k_nonpressure(const __grid_constant__ Params P, StepState *st, const double4 *xrho, const double4 *vel, const double4 *bpos,
                                                      const double4 *bvel, NbrList lf, NbrList lb,
                                                      const double4 *normal, const int *state, double *kappav, int scale_kappav,
                                                      double4 *acc_out, double4 *vel_out, const GhostOut GO, int gate, const VSched S) {
  vsched_prologue(S);
  if (gate == 1 && st->np_done) return;  // graph stepping: the fused pass of the last divergence iteration did this work
  const int nf = st->nf;
  const double h = st->h_step;
  DFR_VB_LOOP(S) {
  const int i = vb_ * 128 + DFR_TID;
  double mag = 0.0;
  (void)nf;
  if (i >= st->own_begin && i < st->own_end) {
    const double4 pi = xrho[i];
    const double4 vi = vel[i];
    d3 a = mk3(P.gx, P.gy, P.gz);
    const double rhoi = pi.w;
    const bool st_on = (P.st_method == 2);
    const bool visc_on = (P.visc_method == 1);
    d3 ni = mk3(0, 0, 0);
    if (st_on) {
      const double4 n4 = normal[i];
      ni = mk3(n4.x, n4.y, n4.z);
    }
    const double h2s = P.support_radius * P.support_radius;
    // (x_j, rho_j) comes as one record (xrho); the normal and the velocity are gathered only when their
    // force is switched on
    for_neighbors4<2>(  // two gather triples in flight (the standalone pass has the registers for it)
       
        lf, i, i,
        [&](int j) {
          Rec3 q;
          q.a = ldg4(xrho + j);
          if (st_on) q.b = ldg4(normal + j);
          if (visc_on) q.c = ldg4(vel + j);
          return q;
        },
        [&](const Rec3 &q, int) {
          const d3 r = mk3(pi.x - q.a.x, pi.y - q.a.y, pi.z - q.a.z);
          const double r2 = dot(r, r);
          double c = 0.0, vdotr = 0.0;
          if (visc_on) {
            c = cubic_grad_coeff(P, r2);
            vdotr = (vi.x - q.c.x) * r.x + (vi.y - q.c.y) * r.y + (vi.z - q.c.z) * r.z;
          }
          nonpressure_pair(P, st_on, visc_on, r, r2, c, rhoi, q.a.w, ni, q.b, vdotr, h2s, a);
        });
    if ((st_on && P.surface_tension_b != 0.0) || (visc_on && P.viscosity_b != 0.0)) {
      // boundary adhesion / boundary viscosity: zero coefficients in every shipped scene; the
      // reaction force of the boundary-viscosity term on dynamic bodies is not carried here.
      for_neighbors4(
          lb, i, 0, [&](int j) { return Rec2{ldg4(bpos + j), ldg4(bvel + j)}; },
          [&](const Rec2 &q, int) {
            const d3 r = mk3(pi.x - q.a.x, pi.y - q.a.y, pi.z - q.a.z);
            const double r2 = dot(r, r);
            if (st_on && P.surface_tension_b != 0.0 && r2 > 1.0e-9)
              a -= (P.surface_tension_b * P.density0 * q.a.w * adhesion_W(P, r2) * rsqrt(r2)) * r;
            if (visc_on && P.viscosity_b != 0.0) {
              const double c = cubic_grad_coeff(P, r2);
              const double vx = (vi.x - q.b.x) * r.x + (vi.y - q.b.y) * r.y + (vi.z - q.b.z) * r.z;
              a += (10.0 * P.viscosity_b * (P.density0 * q.a.w / rhoi) * vx / (r2 + 0.01 * h2s) * c) * r;
            }
          });
    }
    acc_out[i] = make_double4(a.x, a.y, a.z, 0.0);
    const d3 vn = mk3(vi.x + h * a.x, vi.y + h * a.y, vi.z + h * a.z);
    mag = dot(vn, vn);  // CFL uses (vel + accel*h) with the OLD h for every particle (Simulation.cpp:561-566)
    const double4 vo = (state[i] == 0) ? make_double4(vn.x, vn.y, vn.z, 0.0) : vi;
    vel_out[i] = vo;
    ghost_store(GO, i, vo);
    if (scale_kappav) kappav[i] *= h;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mag = fmax(mag, __shfl_xor_sync(DFR_FULL, mag, o));
  // one atomic per warp (a maximum does not depend on the order; no barrier, so the kernel works for any DFR_CTA_VB)
  if ((threadIdx.x & 31) == 0 && mag > 0.0) atomicMax(&st->cfl_max_bits, (unsigned long long)__double_as_longlong(mag));
  }
}

// v += h a, CFL maximum and kappa_v rescale for accelerations that a fused k_rho pass wrote (RHO_X_NONPRESSURE):
// the tail of k_nonpressure as a streaming pass (TimeStepDiffDFSPH.cpp:589-604, Simulation.cpp:542-575)
__global__ void __launch_bounds__(128) k_apply_accel(StepState *st, const double4 *acc, const double4 *vel, const int *state, double *kappav,
                                                      int scale_kappav, double4 *vel_out, const GhostOut GO, int gate) {
  if (gate == 2 && !st->np_done) return;  // graph stepping: k_nonpressure ran instead
  const double h = st->h_step;
  const int i = blockIdx.x * 128 + threadIdx.x;
  double mag = 0.0;
  if (i >= st->own_begin && i < st->own_end) {
    const double4 a = acc[i];
    const double4 vi = vel[i];
    const d3 vn = mk3(vi.x + h * a.x, vi.y + h * a.y, vi.z + h * a.z);
    mag = dot(vn, vn);
    const double4 vo = (state[i] == 0) ? make_double4(vn.x, vn.y, vn.z, 0.0) : vi;
    vel_out[i] = vo;
    ghost_store(GO, i, vo);
    if (scale_kappav) kappav[i] *= h;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mag = fmax(mag, __shfl_xor_sync(DFR_FULL, mag, o));
  __shared__ double wmax[4];
  if ((threadIdx.x & 31) == 0) wmax[threadIdx.x >> 5] = mag;
  __syncthreads();
  if (threadIdx.x == 0) {
    const double m = fmax(fmax(wmax[0], wmax[1]), fmax(wmax[2], wmax[3]));
    atomicMax(&st->cfl_max_bits, (unsigned long long)__double_as_longlong(m));
  }
}

// max |v_b|^2 over the particles of dynamic or animated bodies (Simulation.cpp:577-590)
__global__ void k_cfl_boundary(StepState *st, const double4 *bvel, int dyn_begin, int n_dyn) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  double mag = 0.0;
  if (t < n_dyn) {
    const double4 v = bvel[dyn_begin + t];
    mag = v.x * v.x + v.y * v.y + v.z * v.z;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mag = fmax(mag, __shfl_xor_sync(DFR_FULL, mag, o));
  if ((threadIdx.x & 31) == 0) atomicMax(&st->cfl_max_bits, (unsigned long long)__double_as_longlong(mag));
}

// Simulation::updateTimeStepSize (Simulation.cpp:524-616) + solver-control reset for the pressure solve
__global__ void k_cfl_finish(const __grid_constant__ Params P, StepState *st) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  // close the divergence solve bookkeeping
  st->last_iters_v = st->div_iters;
  st->total_iters_v += st->div_iters;
  if (P.cfl_method == 1 || P.cfl_method == 2) {
    const double maxVel = fmax(0.1, __longlong_as_double((long long)st->cfl_max_bits));
    const double diameter = 2.0 * P.particle_radius;
    double h = P.cfl_factor * 0.4 * (diameter / sqrt(maxVel));
    h = fmin(h, P.cfl_max);
    h = fmax(h, P.cfl_min);
    if (P.cfl_method == 2) {
      double h_old = st->h;
      if (st->last_iters > 10)
        h_old *= 0.9;
      else if (st->last_iters < 5)
        h_old *= 1.1;
      h = fmin(h_old, h);
    }
    st->h = h;
  }
  st->cfl_max_bits = 0ull;
  st->prs_active = 1;
  st->prs_iters = 0;
  st->ticket = 0;
}

// x += h_step v (TimeStepDiffDFSPH.cpp:615-634), kappa *= h^2 (:752-767)
__global__ void k_advect_x(const StepState *st, double4 *pos, const double4 *vel, const int *state, double *kappa, int scale_kappa) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= st->nf) return;
  if (state[i] == 0) {
    const double h = st->h_step;
    double4 p = pos[i];
    const double4 v = vel[i];
    p.x = __dadd_rn(p.x, __dmul_rn(h, v.x));  // mul + add like the reference's x86 code, not an FMA (see k_emit_animate)
    p.y = __dadd_rn(p.y, __dmul_rn(h, v.y));
    p.z = __dadd_rn(p.z, __dmul_rn(h, v.z));
    pos[i] = p;
  }
  if (scale_kappa) kappa[i] *= st->h * st->h;
}

// ---------------------------------------------------------------------------------------------
// Emitter: EmitterSystem::step (EmitterSystem.cpp:54-83) and Emitter::emitParticles (Emitter.cpp:89-227), box
// emitter without particle reuse.  Runs after x += h v and before the time advance (TimeStepDiffDFSPH.cpp:637-645).
// ---------------------------------------------------------------------------------------------
// particles animated by an emitter in the previous step become Active again (EmitterSystem.cpp:64-76)
__global__ void k_emit_release(const StepState *st, int *state) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < st->nf && state[i] == 1) state[i] = 0;
}
__device__ __forceinline__ d3 emitter_velocity(const Params &P, const EmitterDev &e, double t) {
  const d3 dir = mk3(e.rot.a[0], e.rot.a[3], e.rot.a[6]);  // first column of the rotation
  if (t < e.emit_start || t > e.emit_end) return (P.particle_radius * 10.0 * (1.0 / 0.25)) * dir;  // Emitter.cpp:112-113
  return e.velocity * dir;
}
// particles inside the emitter box move with the emit velocity (Emitter.cpp:116-140)
__global__ void k_emit_animate(const __grid_constant__ Params P, const StepState *st, const EmitterDev *emitters, int ei, double4 *pos,
                               double4 *vel, int *state) {
  const EmitterDev &e = emitters[ei];
  const double t = st->time;
  if (!(t >= e.emit_start - 0.25 && t <= e.emit_end)) return;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= st->nf) return;
  const d3 dir = mk3(e.rot.a[0], e.rot.a[3], e.rot.a[6]);
  const d3 ev = emitter_velocity(P, e, t);
  const double diam = 2.0 * P.particle_radius;
  const d3 half = mk3(0.5 * (2.0 * P.support_radius), 0.5 * (e.height * diam + 2.0 * diam), 0.5 * (e.width * diam + 2.0 * diam));
  const d3 c0 = e.x + (0.5 * P.support_radius) * dir;
  double4 p = pos[i];
  const d3 xl = transpose(e.rot) * (mk3(p.x, p.y, p.z) - c0);
  if (fabs(xl.x) < half.x && fabs(xl.y) < half.y && fabs(xl.z) < half.z) {
    const double h = st->h;
    vel[i] = make_double4(ev.x, ev.y, ev.z, 0.0);
    // no FMA contraction: emitted sheets sit on an exact 2r lattice (pairs at distance == support radius), so their
    // positions have to be the bits the reference's mul + add produces
    p.x = __dadd_rn(p.x, __dmul_rn(h, ev.x));
    p.y = __dadd_rn(p.y, __dmul_rn(h, ev.y));
    p.z = __dadd_rn(p.z, __dmul_rn(h, ev.z));
    pos[i] = p;
    state[i] = 1;
  }
}
// appends width x height particles when the emit time has come (Emitter.cpp:142-227); one block
__global__ void k_emit_spawn(const __grid_constant__ Params P, StepState *st, EmitterDev *emitters, int ei, int capacity, double4 *pos,
                             double4 *vel, double *kappa, double *kappav, int *pid, int *state) {
  EmitterDev &e = emitters[ei];
  const double t = st->time;
  if (t < e.next_emit_time || t > e.emit_end) return;
  const int nf = st->nf;
  const double diam = 2.0 * P.particle_radius;
  const d3 ev = emitter_velocity(P, e, t);
  const d3 axisH = mk3(e.rot.a[1], e.rot.a[4], e.rot.a[7]);
  const d3 axisW = mk3(e.rot.a[2], e.rot.a[5], e.rot.a[8]);
  const double startX = -0.5 * (e.width - 1) * diam;
  const double startZ = -0.5 * (e.height - 1) * diam;
  const double dt = t - e.next_emit_time + st->h;
  // every product and sum below is rounded on its own (no FMA contraction), in the reference's order (Emitter.cpp:171):
  // a sheet is an exact 2r lattice with pairs at distance == support radius, and Akinci-2013's normal term is not
  // kernel-weighted, so one differing bit in an emitted position can flip such a pair in or out of the lists and change
  // the surface-tension force by O(1)
  const d3 offset = mk3(__dadd_rn(e.x.x, __dmul_rn(dt, ev.x)), __dadd_rn(e.x.y, __dmul_rn(dt, ev.y)), __dadd_rn(e.x.z, __dmul_rn(dt, ev.z)));
  if (nf < capacity) {
    const int total = e.width * e.height;
    for (int k = threadIdx.x; k < total; k += blockDim.x) {
      const int i = k / e.height, j = k % e.height;
      const int index = nf + k;
      if (index < capacity) {
        const double sw = __dadd_rn(__dmul_rn((double)i, diam), startX), sh = __dadd_rn(__dmul_rn((double)j, diam), startZ);
        const d3 x = mk3(__dadd_rn(__dadd_rn(__dmul_rn(sw, axisW.x), __dmul_rn(sh, axisH.x)), offset.x),
                         __dadd_rn(__dadd_rn(__dmul_rn(sw, axisW.y), __dmul_rn(sh, axisH.y)), offset.y),
                         __dadd_rn(__dadd_rn(__dmul_rn(sw, axisW.z), __dmul_rn(sh, axisH.z)), offset.z));
        pos[index] = make_double4(x.x, x.y, x.z, 0.0);
        vel[index] = make_double4(ev.x, ev.y, ev.z, 0.0);
        state[index] = 1;
        kappa[index] = 0.0;  // SimulationDataDiffDFSPH::emittedParticles (:173-183)
        kappav[index] = 0.0;
        pid[index] = index;
      }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    if (nf < capacity) st->nf = min(nf + e.width * e.height, capacity);
    e.next_emit_time += diam / e.velocity;
    e.emit_counter += 1;
  }
}

}  // namespace dfr

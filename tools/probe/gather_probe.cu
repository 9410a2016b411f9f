// Microbenchmark (tuning probe, not part of the product): cost of a warp-wide gather of 32-byte records from an
// L1-resident table as a function of the lane -> record pattern and of the load width, from global memory (L1) and
// from shared memory.  Prints SM cycles per warp-gather (all SMs busy, many warps per SM).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather_probe gather_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define NREC_SMALL 2048  // 64 KB table: L1 resident
__constant__ unsigned int c_nrec = NREC_SMALL;  // records in the table (power of two)
#define NREC c_nrec
__device__ __forceinline__ unsigned int hash32(unsigned int x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}
// lane -> record index for pattern P; s is an LCG state seeded per lane / per lane group / per warp (see seed_for)
template <int P>
__device__ __forceinline__ unsigned int seed_for(int lane, unsigned int wid) {
  switch (P) {
    case 4: case 9: return hash32(wid * 64u + lane);                  // per lane
    case 3: return hash32(wid * 64u + (lane >> 1));                   // per pair
    case 2: case 10: return hash32(wid * 64u + (lane >> 2));          // per 4-lane group
    case 11: return hash32(wid * 64u + (lane >> 3));                  // per 8-lane group
    default: return hash32(wid * 64u + 63u);                          // per warp
  }
}
template <int P>
__device__ __forceinline__ int pattern(int lane, unsigned int &s) {
  s = s * 1664525u + 1013904223u;
  const unsigned int b = s >> 9;
  switch (P) {
    case 0: return ((b & ~31u) + lane) % NREC;                       // 32 consecutive records, line aligned
    case 1: return ((b & ~31u) + lane + 1) % NREC;                   // consecutive, misaligned by one record
    case 2: return ((b & ~3u) + (lane & 3)) % NREC;                  // aligned 4-lane groups, groups random
    case 3: return ((b & ~1u) + (lane & 1)) % NREC;                  // aligned pairs, pairs random
    case 4: return b % NREC;                                         // every lane random
    case 5: return b % NREC;                                         // all lanes the same record
    case 6: return (b + (lane >> 2) + (lane & 3)) % NREC;            // 8 quads, windows of 4 records shifted by one
    case 7: return (b + 2 * lane) % NREC;                            // stride 2 records
    case 8: return (b + 4 * lane) % NREC;                            // stride 4 records: one line per lane
    case 9: return ((hash32(s ^ 0x9e3779b9u) & ~31u) + lane + (b & 3)) % NREC;  // placeholder, replaced below
    case 10: return (b + (lane & 3)) % NREC;                         // unaligned 4-lane groups, groups random
    case 11: return ((b & ~7u) + (lane & 7)) % NREC;                 // aligned 8-lane groups (2 lines), random
    case 12: return (b + (lane >> 2) * 3 + (lane & 3)) % NREC;       // quads, windows shifted by 3 (overlap 1)
    case 13: return (b + (lane >> 2) * 8 + (lane & 3)) % NREC;       // quads, one line each (if aligned), every 2nd line
  }
  return 0;
}

template <int P, int W /*0: 256-bit, 1: 2x128-bit, 2: shared 2x LDS.128, 3: 4x64-bit global*/>
__global__ void __launch_bounds__(256) k_probe(const double4 *tab, double *out, int iters) {
  extern __shared__ double4 stab[];
  if (W == 2) {
    for (int i = threadIdx.x; i < NREC_SMALL; i += blockDim.x) stab[i] = tab[i];
    __syncthreads();
  }
  const int lane = threadIdx.x & 31;
  const unsigned int wid = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  double acc = 0.0;
  unsigned int st = seed_for<P>(lane, wid);
  unsigned int sw = hash32(wid * 64u + 63u);  // warp-uniform stream (pattern 9 base)
  for (int t = 0; t < iters; t++) {
    int j = pattern<P>(lane, st);
    if (P == 9) {  // consecutive + per-lane jitter 0..3 around a warp-uniform base
      sw = sw * 1664525u + 1013904223u;
      j = ((sw >> 9) + lane + ((st >> 20) & 3)) % NREC;
    }
    double4 v;
    if (W == 0) {
      asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v.x), "=d"(v.y), "=d"(v.z), "=d"(v.w) : "l"(tab + j));
    } else if (W == 1) {
      const double2 *p = reinterpret_cast<const double2 *>(tab + j);
      double2 a, b;
      asm volatile("ld.global.nc.v2.f64 {%0,%1}, [%2];" : "=d"(a.x), "=d"(a.y) : "l"(p));
      asm volatile("ld.global.nc.v2.f64 {%0,%1}, [%2];" : "=d"(b.x), "=d"(b.y) : "l"(p + 1));
      v = make_double4(a.x, a.y, b.x, b.y);
    } else if (W == 2) {
      const double2 *p = reinterpret_cast<const double2 *>(stab + j);
      const double2 a = p[0], b = p[1];
      v = make_double4(a.x, a.y, b.x, b.y);
    } else {
      const double *p = reinterpret_cast<const double *>(tab + j);
      double a, b, c, d;
      asm volatile("ld.global.nc.f64 %0, [%1];" : "=d"(a) : "l"(p));
      asm volatile("ld.global.nc.f64 %0, [%1];" : "=d"(b) : "l"(p + 1));
      asm volatile("ld.global.nc.f64 %0, [%1];" : "=d"(c) : "l"(p + 2));
      asm volatile("ld.global.nc.f64 %0, [%1];" : "=d"(d) : "l"(p + 3));
      v = make_double4(a, b, c, d);
    }
    acc += v.x + v.w;
  }
  if (acc == 123.456) out[0] = acc;
}

template <int P, int W>
void run(const double4 *tab, double *out, int nsm, double mhz) {
  const int iters = 2000, blocks = nsm * 8;
  const size_t sh = (W == 2) ? NREC_SMALL * sizeof(double4) : 0;
  if (W == 2) cudaFuncSetAttribute(k_probe<P, W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh);
  const int nb = (W == 2) ? nsm * 3 : blocks;
  k_probe<P, W><<<nb, 256, sh>>>(tab, out, 100);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k_probe<P, W><<<nb, 256, sh>>>(tab, out, iters);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  const double gathers_per_sm = (double)nb * 8 * iters / nsm;
  const double cyc = ms * 1e-3 * mhz * 1e6;
  printf("  W%d: %6.2f cyc/gather", W, cyc / gathers_per_sm);
}
static bool g_small = true;
template <int P>
void row(const char *name, const double4 *tab, double *out, int nsm, double mhz) {
  printf("P%-2d %-52s", P, name);
  run<P, 0>(tab, out, nsm, mhz);
  run<P, 1>(tab, out, nsm, mhz);
  run<P, 3>(tab, out, nsm, mhz);
  if (g_small) run<P, 2>(tab, out, nsm, mhz);
  printf("\n");
}
int main(int argc, char **argv) {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  int khz = 0;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  const double mhz = khz / 1000.0;
  double4 *tab;
  double *out;
  const unsigned int nrec = argc > 1 ? (unsigned int)atoi(argv[1]) : NREC_SMALL;  // e.g. 2097152 = 64 MB: L2 resident, L1 misses
  cudaMemcpyToSymbol(c_nrec, &nrec, sizeof(nrec));
  g_small = (nrec == NREC_SMALL);
  cudaMalloc(&tab, (size_t)nrec * sizeof(double4));
  cudaMemset(tab, 0, (size_t)nrec * sizeof(double4));
  printf("table: %u records (%.1f MB)\n", nrec, nrec * 32.0 / 1e6);
  cudaMalloc(&out, 8);
  printf("SMs %d clock %.0f MHz; cycles per warp-wide gather of 32-byte records (W0 LDG.256, W1 2xLDG.128, W3 4xLDG.64, W2 shared 2xLDS.128)\n",
         p.multiProcessorCount, mhz);
  row<0>("32 consecutive records, aligned", tab, out, p.multiProcessorCount, mhz);
  row<1>("32 consecutive, misaligned by one", tab, out, p.multiProcessorCount, mhz);
  row<2>("aligned 4-lane groups (one line each), random", tab, out, p.multiProcessorCount, mhz);
  row<10>("unaligned 4-lane groups, random", tab, out, p.multiProcessorCount, mhz);
  row<11>("aligned 8-lane groups (two lines each), random", tab, out, p.multiProcessorCount, mhz);
  row<3>("aligned pairs, random", tab, out, p.multiProcessorCount, mhz);
  row<4>("every lane random", tab, out, p.multiProcessorCount, mhz);
  row<5>("all lanes same record", tab, out, p.multiProcessorCount, mhz);
  row<6>("8 quads, 4-record windows shifted by one", tab, out, p.multiProcessorCount, mhz);
  row<12>("8 quads, windows shifted by three", tab, out, p.multiProcessorCount, mhz);
  row<13>("8 quads, one line each, consecutive lines x2", tab, out, p.multiProcessorCount, mhz);
  row<7>("stride 2 records", tab, out, p.multiProcessorCount, mhz);
  row<8>("stride 4 records (one line per lane)", tab, out, p.multiProcessorCount, mhz);
  row<9>("consecutive + jitter 0..3", tab, out, p.multiProcessorCount, mhz);
  return 0;
}

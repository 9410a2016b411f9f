// prints the %smid values a persistent grid sees (tuning probe, not part of the product)
#include <cstdio>
#include <cuda_runtime.h>
#include <set>
#include <vector>
__global__ void k(unsigned int *out, unsigned int *nsm) {
  unsigned int s, n;
  asm("mov.u32 %0, %%smid;" : "=r"(s));
  asm("mov.u32 %0, %%nsmid;" : "=r"(n));
  if (threadIdx.x == 0) out[blockIdx.x] = s;
  if (blockIdx.x == 0 && threadIdx.x == 0) *nsm = n;
  // spin a little so that all CTAs are co-resident
  long long t0 = clock64();
  while (clock64() - t0 < 200000) {}
}
int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  const int n = p.multiProcessorCount * 8;
  unsigned int *d, *dn;
  cudaMalloc(&d, n * 4);
  cudaMalloc(&dn, 4);
  k<<<n, 128>>>(d, dn);
  std::vector<unsigned int> h(n);
  unsigned int nsm;
  cudaMemcpy(h.data(), d, n * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(&nsm, dn, 4, cudaMemcpyDeviceToHost);
  std::set<unsigned int> ids(h.begin(), h.end());
  std::vector<int> cnt(512, 0);
  for (auto v : h) cnt[v]++;
  int mn = 1 << 30, mx = 0;
  for (auto v : ids) { mn = std::min(mn, cnt[v]); mx = std::max(mx, cnt[v]); }
  printf("multiProcessorCount %d nsmid %u distinct smids %zu min id %u max id %u ctas/sm min %d max %d\n", p.multiProcessorCount, nsm, ids.size(), *ids.begin(), *ids.rbegin(), mn, mx);
  printf("first 40 block->smid:");
  for (int i = 0; i < 40; i++) printf(" %u", h[i]);
  printf("\n");
  return 0;
}

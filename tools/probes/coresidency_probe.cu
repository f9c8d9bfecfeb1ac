// Can a small kernel run on an SM BESIDE a CTA that holds ~all of its shared memory?  (B200 probe for
// the flag-mode schedule: the tensor-core kernel waits for dictionary tiles that a normalise kernel
// on another stream produces.)  Kernel A: one CTA per SM with `smem_a` bytes of dynamic shared memory,
// spins until a flag is set (or a timeout).  Kernel B (launched afterwards on another stream): 256
// threads, 64 bytes of static shared memory, sets the flag.  Reported per combination: did A see the
// flag before its timeout, i.e. did B get onto the device while A occupied every SM?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o coresidency_probe coresidency_probe.cu
#include <cuda_runtime.h>
#include <cstdio>

__global__ void __launch_bounds__(192, 1) spin_kernel(volatile unsigned* flag, unsigned* seen, long long timeout) {
  extern __shared__ unsigned char smem[];
  if (threadIdx.x == 0) {
    smem[0] = 1;
    const long long t0 = clock64();
    unsigned ok = 0;
    while (clock64() - t0 < timeout) {
      if (*flag) { ok = 1; break; }
      __nanosleep(200);
    }
    if (ok) atomicAdd(seen, 1u);
  }
}

__global__ void __launch_bounds__(256) set_kernel(unsigned* flag, unsigned* ran) {
  __shared__ double red[8];
  if (threadIdx.x < 8) red[threadIdx.x] = threadIdx.x;
  __syncthreads();
  if (threadIdx.x == 0) {
    atomicAdd(ran, 1u);
    if (red[3] == 3.0) { __threadfence(); *flag = 1u; }
  }
}

int main() {
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  const int sms = prop.multiProcessorCount;
  printf("device %s, %d SMs, %zu B shared memory per SM, %zu B max per block (opt-in)\n", prop.name, sms,
         prop.sharedMemPerMultiprocessor, prop.sharedMemPerBlockOptin);
  unsigned *flag, *seen, *ran;
  cudaMalloc(&flag, 4); cudaMalloc(&seen, 4); cudaMalloc(&ran, 4);
  cudaStream_t sa, sb;
  int lo, hi;
  cudaDeviceGetStreamPriorityRange(&lo, &hi);
  cudaStreamCreateWithPriority(&sa, cudaStreamNonBlocking, hi);
  cudaStreamCreateWithPriority(&sb, cudaStreamNonBlocking, lo);
  const int smem_sizes[] = {230544, 197776, 165008, 100000};
  const int carve_b[] = {-1, 100, 0};
  for (int smem_a : smem_sizes) {
    for (int cb : carve_b) {
      cudaFuncSetAttribute(spin_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_a);
      cudaFuncSetAttribute(spin_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
      cudaFuncSetAttribute(set_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cb);
      cudaMemset(flag, 0, 4); cudaMemset(seen, 0, 4); cudaMemset(ran, 0, 4);
      cudaDeviceSynchronize();
      const long long timeout = 400000000LL;  // ~0.2 s
      spin_kernel<<<sms, 192, smem_a, sa>>>(flag, seen, timeout);
      // give A time to occupy every SM, then launch B
      cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, sa);
      for (volatile int i = 0; i < 2000000; ++i) {}
      set_kernel<<<4 * sms, 256, 0, sb>>>(flag, ran);
      cudaError_t err = cudaDeviceSynchronize();
      unsigned h_seen = 0, h_ran = 0;
      cudaMemcpy(&h_seen, seen, 4, cudaMemcpyDeviceToHost);
      cudaMemcpy(&h_ran, ran, 4, cudaMemcpyDeviceToHost);
      printf("A smem %6d B (free per SM %6lld B)  B carveout %3d : A CTAs that saw the flag %3u / %d  (B CTAs run %u) %s\n",
             smem_a, (long long)prop.sharedMemPerMultiprocessor - smem_a - 1024, cb, h_seen, sms, h_ran,
             err == cudaSuccess ? "" : cudaGetErrorString(err));
    }
  }
  return 0;
}

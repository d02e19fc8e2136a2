// Probe of programmatic dependent launch semantics on this GPU / driver: a persistent primary that publishes per-task flags, a
// dependent launch that consumes them while the primary is still running, and a third kernel dependent on the second.
// Prints, per configuration, whether the consumers saw every flag (bounded spins) and how long the chain took.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ int ld_acquire(const int* p) { int v; asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void st_release(int* p, int v) { asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }

__global__ void producer(int* counter, int ntasks, int* flags, int* data, int epoch, int work, int trigger_mode) {
  const int lane = threadIdx.x & 31;
  if (trigger_mode == 1) pdl_trigger();
  for (;;) {
    int t = 0;
    if (lane == 0) t = atomicAdd(counter, 1);
    t = __shfl_sync(0xffffffffu, t, 0);
    if (t >= ntasks) break;
    if (trigger_mode == 2 && t >= ntasks - (int)(gridDim.x * blockDim.x / 32)) pdl_trigger();
    long long t0 = clock64();
    while (clock64() - t0 < (long long)work * (1 + (t % 7))) {}
    data[t * 32 + lane] = epoch + t;
    __threadfence();
    __syncwarp();
    if (lane == 0) st_release(&flags[t], epoch);
  }
  pdl_trigger();
}
__global__ void consumer(int* counter, int ntasks, const int* wait, int* flags, const int* in, int* out, int epoch, int* stuck, int wait_at_end) {
  const int lane = threadIdx.x & 31;
  pdl_trigger();
  for (;;) {
    int t = 0;
    if (lane == 0) t = atomicAdd(counter, 1);
    t = __shfl_sync(0xffffffffu, t, 0);
    if (t >= ntasks) break;
    if (lane == 0) {
      unsigned spins = 0;
      while (ld_acquire(&wait[t]) != epoch) { __nanosleep(400); if (++spins > (1u << 21)) { atomicAdd(stuck, 1); break; } }
    }
    __syncwarp();
    out[t * 32 + lane] = __ldcg(&in[t * 32 + lane]) + 1;
    __threadfence();
    __syncwarp();
    if (lane == 0) st_release(&flags[t], epoch);
  }
  if (wait_at_end) pdl_wait();
}

static void launch(void (*k)(int*, int, int*, int*, int, int, int), int grid, cudaStream_t st, int* c, int n, int* f, int* d, int e, int w, int tm) {
  k<<<grid, 128, 0, st>>>(c, n, f, d, e, w, tm);
}
int main(int argc, char** argv) {
  const int ntasks = argc > 1 ? atoi(argv[1]) : 20000;
  int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  int *counters, *f1, *f2, *f3, *d1, *d2, *d3, *stuck;
  cudaMalloc(&counters, 64 * 4); cudaMalloc(&f1, ntasks * 4); cudaMalloc(&f2, ntasks * 4); cudaMalloc(&f3, ntasks * 4);
  cudaMalloc(&d1, ntasks * 128); cudaMalloc(&d2, ntasks * 128); cudaMalloc(&d3, ntasks * 128); cudaMalloc(&stuck, 4);
  cudaMemset(f1, 0, ntasks * 4); cudaMemset(f2, 0, ntasks * 4); cudaMemset(f3, 0, ntasks * 4);
  cudaStream_t st; cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  int epoch = 0;
  for (int cfg = 0; cfg < 8; cfg++) {
    const int pdl = cfg & 1, trigger_mode = (cfg >> 1) & 3;          // trigger: 0 at exit only, 1 at start, 2 in the last round
    if (trigger_mode == 3) continue;
    for (int wait_end = 0; wait_end < 2; wait_end++) {
      if (!pdl && wait_end) continue;
      epoch++;
      cudaMemsetAsync(counters, 0, 64 * 4, st); cudaMemsetAsync(stuck, 0, 4, st);
      cudaEventRecord(e0, st);
      launch(producer, sms * 16, st, counters, ntasks, f1, d1, epoch, 20000, trigger_mode);
      for (int stage = 0; stage < 2; stage++) {
        cudaLaunchConfig_t lc = {};
        lc.gridDim = dim3(sms * 16); lc.blockDim = dim3(128); lc.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
        lc.attrs = at; lc.numAttrs = pdl ? 1 : 0;
        int* cn = counters + 1 + stage; const int* wt = stage ? f2 : f1; int* fo = stage ? f3 : f2; const int* in = stage ? d2 : d1; int* out = stage ? d3 : d2;
        cudaLaunchKernelEx(&lc, consumer, cn, ntasks, wt, fo, in, out, epoch, stuck, wait_end);
      }
      cudaEventRecord(e1, st);
      cudaError_t err = cudaStreamSynchronize(st);
      float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
      int h_stuck = -1, last = 0; cudaMemcpy(&h_stuck, stuck, 4, cudaMemcpyDeviceToHost); cudaMemcpy(&last, d3 + (ntasks - 1) * 32, 4, cudaMemcpyDeviceToHost);
      printf("pdl=%d trigger=%d wait_at_end=%d: %s  %.3f ms  stuck=%d  check=%s\n", pdl, trigger_mode, wait_end, cudaGetErrorString(err), ms, h_stuck,
             last == epoch + ntasks - 1 + 2 ? "ok" : "BAD");
      fflush(stdout);
    }
  }
  return 0;
}

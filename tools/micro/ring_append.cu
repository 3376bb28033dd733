// Micro-benchmark: appending one 47-float and one 73-float frame per env to both mirrored slots of the history rings
// (rows of 2*66*47 and 2*3*73 floats per env, slot = step % H) from frames staged in shared memory, 32 envs per CTA:
//   A  flat scalar loop: the CTA's warps stride over (env, k), two 4-byte stores per element (rounds 1-2)
//   B  one warp per run: 128-bit stores between the 16-byte boundaries, scalars for head / tail
//   C  one THREAD per run: the 16-byte aligned middle as ONE bulk copy shared -> global (cp.async.bulk), head / tail
//      scalars by the same thread; frames staged at the smem offset that matches the run's global alignment
// nvcc -arch=sm_100a -O3 -o tools/micro/ring_append tools/micro/ring_append.cu && tools/micro/ring_append
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
constexpr int K = 47, P = 73, H = 66, CH = 3, TB = 32;
constexpr size_t OROW = 2 * H * K, PROW = 2 * CH * P;
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int MODE, int NT>
__global__ void __launch_bounds__(NT) append(float* obs_ring, float* priv_ring, int N, int step) {
  // staged copies: [dst 0/1][env][row of 80 floats], frame at float offset = alignment of its destination
  __shared__ __align__(16) float s_obs[2][TB][52];
  __shared__ __align__(16) float s_priv[2][TB][80];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, e0 = blockIdx.x * TB;
  const int hs = step % H, cs = step % CH;
  const int n_tile = min(TB, N - e0);
  // alignment (floats mod 4) of the two obs runs is uniform; of the priv runs it depends on the env's parity
  for (int i = tid; i < n_tile * K; i += NT) {
    const int en = i / K, k = i - en * K;
    const float v = (float)(e0 + en) + 0.001f * k;
    for (int d = 0; d < 2; ++d) {
      const size_t off = (size_t)(e0 + en) * OROW + (size_t)(hs + d * H) * K;
      s_obs[d][en][(MODE == 2 ? (off & 3) : 0) + k] = v;
    }
  }
  for (int i = tid; i < n_tile * P; i += NT) {
    const int en = i / P, k = i - en * P;
    const float v = (float)(e0 + en) + 0.001f * k;
    for (int d = 0; d < 2; ++d) {
      const size_t off = (size_t)(e0 + en) * PROW + (size_t)(cs + d * CH) * P;
      s_priv[d][en][(MODE == 2 ? (off & 3) : 0) + k] = v;
    }
  }
  __syncthreads();
  if (MODE == 0) {
    float* ob = obs_ring + (size_t)e0 * OROW + (size_t)hs * K;
    float* pb = priv_ring + (size_t)e0 * PROW + (size_t)cs * P;
#pragma unroll 1
    for (int i = tid; i < n_tile * K; i += NT) {
      const int en = i / K, k = i - en * K;
      const float v = s_obs[0][en][k];
      float* dst = ob + (uint32_t)en * (uint32_t)OROW + k;
      dst[0] = v; dst[H * K] = v;
    }
#pragma unroll 1
    for (int i = tid; i < n_tile * P; i += NT) {
      const int en = i / P, k = i - en * P;
      const float v = s_priv[0][en][k];
      float* dst = pb + (uint32_t)en * (uint32_t)PROW + k;
      dst[0] = v; dst[CH * P] = v;
    }
  } else if (MODE == 3) {
    // slot-major layout (2H, N, K): what a streaming write of the same bytes costs
    float* ob = obs_ring + ((size_t)hs * N + e0) * K;
    float* pb = priv_ring + ((size_t)cs * N + e0) * P;
    for (int i = tid; i < n_tile * K; i += NT) { const float v = s_obs[0][i / K][i % K]; ob[i] = v; ob[(size_t)H * N * K + i] = v; }
    for (int i = tid; i < n_tile * P; i += NT) { const float v = s_priv[0][i / P][i % P]; pb[i] = v; pb[(size_t)CH * N * P + i] = v; }
  } else if (MODE == 1) {
    auto run = [&](float* dst, const float* src, int len) {
      const int head = (int)((4u - ((uint32_t)(reinterpret_cast<uintptr_t>(dst) >> 2) & 3u)) & 3u);
      const int nv = (len - head) >> 2, tail0 = head + 4 * nv, ns = head + (len - tail0);
      for (int v = lane; v < nv; v += 32) {
        const int k = head + 4 * v;
        *reinterpret_cast<float4*>(dst + k) = make_float4(src[k], src[k + 1], src[k + 2], src[k + 3]);
      }
      const int sidx = lane - (32 - ns);
      if (sidx >= 0) { const int k = sidx < head ? sidx : tail0 + (sidx - head); dst[k] = src[k]; }
    };
    for (int en = warp; en < n_tile; en += NT / 32) {
      float* d0 = obs_ring + (size_t)(e0 + en) * OROW + (size_t)hs * K;
      run(d0, s_obs[0][en], K); run(d0 + H * K, s_obs[0][en], K);
      float* d1 = priv_ring + (size_t)(e0 + en) * PROW + (size_t)cs * P;
      run(d1, s_priv[0][en], P); run(d1 + CH * P, s_priv[0][en], P);
    }
  } else {
    // one thread per run: 4 runs per env
    for (int r = tid; r < n_tile * 4; r += NT) {
      const int en = r >> 2, kind = r & 3, d = kind & 1;
      const bool is_obs = kind < 2;
      const int len = is_obs ? K : P;
      float* dst = is_obs ? obs_ring + (size_t)(e0 + en) * OROW + (size_t)(hs + d * H) * K
                          : priv_ring + (size_t)(e0 + en) * PROW + (size_t)(cs + d * CH) * P;
      const int a = (int)((reinterpret_cast<uintptr_t>(dst) >> 2) & 3u);
      const float* src = (is_obs ? s_obs[d][en] : s_priv[d][en]) + a;       // src[k] <-> dst[k], same alignment mod 16 bytes
      const int head = (4 - a) & 3, nv = (len - head) >> 2, tail0 = head + 4 * nv;
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + head), "r"(smem_u32(src + head)),
                   "r"(nv * 16) : "memory");
      for (int k = 0; k < head; ++k) dst[k] = src[k];
      for (int k = tail0; k < len; ++k) dst[k] = src[k];
    }
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  }
}

__global__ void read_flush(const float4* src, size_t n, float* sink) {
  float acc = 0.f;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) { const float4 v = src[i]; acc += v.x + v.y + v.z + v.w; }
  if (acc == 123.456f) *sink = acc;
}
static bool g_clean = false;
template <int MODE, int NT>
static float run(float* o, float* p, int N, char* flush, size_t fb, int reps) {
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  float tot = 0;
  for (int i = 0; i < reps + 3; ++i) {
    cudaMemsetAsync(flush, i, fb);
    if (g_clean) read_flush<<<148 * 8, 256>>>(reinterpret_cast<const float4*>(flush) + (fb / 32), fb / 32, reinterpret_cast<float*>(flush));
    cudaEventRecord(a);
    append<MODE, NT><<<(N + TB - 1) / TB, NT>>>(o, p, N, 7 + i);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    if (i >= 3) tot += ms;
  }
  return tot / reps * 1e3f;
}

int main() {
  for (int N : {8192, 65536}) {
    float *o, *p, *o2, *p2;
    cudaMalloc(&o, N * OROW * 4); cudaMalloc(&p, N * PROW * 4);
    cudaMalloc(&o2, N * OROW * 4); cudaMalloc(&p2, N * PROW * 4);
    char* flush; const size_t fb = 256u << 20;
    cudaMalloc(&flush, fb);
    // correctness: C against A
    cudaMemset(o, 0, N * OROW * 4); cudaMemset(o2, 0, N * OROW * 4); cudaMemset(p, 0, N * PROW * 4); cudaMemset(p2, 0, N * PROW * 4);
    for (int s = 0; s < 70; ++s) {
      append<0, 128><<<(N + TB - 1) / TB, 128>>>(o, p, N, s);
      if (s & 1) append<2, 128><<<(N + TB - 1) / TB, 128>>>(o2, p2, N, s); else append<1, 128><<<(N + TB - 1) / TB, 128>>>(o2, p2, N, s);
    }
    cudaDeviceSynchronize();
    {
      const size_t n = 1000 * OROW;
      float* h1 = (float*)malloc(n * 4); float* h2 = (float*)malloc(n * 4);
      cudaMemcpy(h1, o, n * 4, cudaMemcpyDeviceToHost); cudaMemcpy(h2, o2, n * 4, cudaMemcpyDeviceToHost);
      size_t bad = 0; for (size_t i = 0; i < n; ++i) bad += h1[i] != h2[i];
      const size_t m = 1000 * PROW;
      cudaMemcpy(h1, p, m * 4, cudaMemcpyDeviceToHost); cudaMemcpy(h2, p2, m * 4, cudaMemcpyDeviceToHost);
      for (size_t i = 0; i < m; ++i) bad += h1[i] != h2[i];
      printf("N=%d mismatches B/C vs A: %zu  (%s)\n", N, bad, cudaGetErrorString(cudaGetLastError()));
      free(h1); free(h2);
    }
    printf("N=%d  A flat scalar 128thr: %.2f us   256thr: %.2f us\n", N, run<0, 128>(o, p, N, flush, fb, 20), run<0, 256>(o, p, N, flush, fb, 20));
    printf("N=%d  B warp/run    128thr: %.2f us   256thr: %.2f us\n", N, run<1, 128>(o, p, N, flush, fb, 20), run<1, 256>(o, p, N, flush, fb, 20));
    printf("N=%d  C bulk/thread 128thr: %.2f us   256thr: %.2f us\n", N, run<2, 128>(o, p, N, flush, fb, 20), run<2, 256>(o, p, N, flush, fb, 20));
    printf("N=%d  D slot-major  128thr: %.2f us\n", N, run<3, 128>(o, p, N, flush, fb, 20));
    g_clean = true;
    printf("N=%d  clean L2 (write then read flush): A %.2f us  B %.2f us  C %.2f us  D %.2f us\n", N, run<0, 128>(o, p, N, flush, fb, 20),
           run<1, 128>(o, p, N, flush, fb, 20), run<2, 128>(o, p, N, flush, fb, 20), run<3, 128>(o, p, N, flush, fb, 20));
    g_clean = false;
    cudaFree(o); cudaFree(p); cudaFree(o2); cudaFree(p2); cudaFree(flush);
  }
  return 0;
}

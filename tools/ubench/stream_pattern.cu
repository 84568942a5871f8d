// A/B of the frame->voxel kernel's memory pattern WITHOUT its arithmetic (B200, sm_100a): what the access pattern itself
// can stream, and whether staging the frame tiles through shared memory with bulk async copies (cp.async.bulk + mbarrier,
// SASS UBLKCP) moves that number.
//   pattern  : grid (tiles, clips); CTA = 512 lanes x 4 pixels; for each of 120 intervals a lane reads one 32-bit word of
//              the next uint8 frame (stride H*W between frames) and writes one 128-bit float4 of the voxel plane.
//   ldg      : the word comes from ld.global.nc.L1::no_allocate, two trips of 4 frames in flight per lane (the kernel's form);
//   bulk     : thread 0 of the CTA keeps a ring of DEPTH 2 KB frame tiles in flight with cp.async.bulk.shared::cluster.global
//              completing on per-slot mbarriers; every warp waits for the slot, reads its word with LDS.32 and one lane per
//              warp arrives on the slot's "empty" barrier; the producer refills a slot when all 16 warps have released it.
//   copy     : plain float4 copy of the same number of bytes (the STREAM-style denominator).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o stream_pattern stream_pattern.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int THREADS = 512, NFRAMES = 121, H = 480, W = 640, CLIPS = 32, DEPTH = 8;
constexpr int HW = H * W;

__device__ __forceinline__ uint32_t ld_stream_u32(const void* p) {
  uint32_t v;
  asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ void st_stream_f32x4(float* p, float a, float b, float c, float d) {
  asm volatile("st.global.cs.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void emit(float* vox, uint32_t w) {      // cheapest possible use of every byte
  st_stream_f32x4(vox, __uint_as_float((w & 0xffu) << 23), __uint_as_float((w & 0xff00u) << 15), __uint_as_float((w & 0xff0000u) << 7),
                  __uint_as_float((w >> 24) << 23));
}

__global__ void __launch_bounds__(THREADS, 2) pattern_ldg(const uint8_t* frames, float* voxel) {
  const int64_t pix0 = (static_cast<int64_t>(blockIdx.x) * THREADS + threadIdx.x) * 4;
  const uint8_t* fr = frames + static_cast<int64_t>(blockIdx.y) * NFRAMES * HW + pix0;
  float* vox = voxel + static_cast<int64_t>(blockIdx.y) * (NFRAMES - 1) * HW + pix0;
  uint32_t cur[4], nxt[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) cur[u] = ld_stream_u32(fr + static_cast<int64_t>(1 + u) * HW);
  for (int t = 0; t < (NFRAMES - 1) / 4; ++t) {
    if (t + 1 < (NFRAMES - 1) / 4) {
#pragma unroll
      for (int u = 0; u < 4; ++u) nxt[u] = ld_stream_u32(fr + static_cast<int64_t>(5 + 4 * t + u) * HW);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      emit(vox, cur[u]);
      vox += HW;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) cur[u] = nxt[u];
  }
}

__device__ __forceinline__ void mbar_init(uint32_t bar, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, int bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n .reg .pred p;\n W: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra D;\n bra W;\n D:\n}" ::"r"(bar), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, int bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes),
               "r"(bar)
               : "memory");
}

__global__ void __launch_bounds__(THREADS, 2) pattern_bulk(const uint8_t* frames, float* voxel) {
  __shared__ __align__(128) uint8_t ring[DEPTH][THREADS * 4];
  __shared__ __align__(8) uint64_t full_b[DEPTH], empty_b[DEPTH];
  const int64_t tile0 = static_cast<int64_t>(blockIdx.x) * THREADS * 4;
  const uint8_t* fr = frames + static_cast<int64_t>(blockIdx.y) * NFRAMES * HW + tile0;
  float* vox = voxel + static_cast<int64_t>(blockIdx.y) * (NFRAMES - 1) * HW + tile0 + threadIdx.x * 4;
  const uint32_t ring_a = static_cast<uint32_t>(__cvta_generic_to_shared(&ring[0][0]));
  const uint32_t full_a = static_cast<uint32_t>(__cvta_generic_to_shared(&full_b[0])), empty_a = static_cast<uint32_t>(__cvta_generic_to_shared(&empty_b[0]));
  if (threadIdx.x == 0) {
    for (int s = 0; s < DEPTH; ++s) {
      mbar_init(full_a + 8 * s, 1);
      mbar_init(empty_a + 8 * s, THREADS / 32);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  constexpr int M = NFRAMES - 1;
  if (threadIdx.x == 0) {                         // prologue: fill the ring
    for (int n = 0; n < DEPTH && n < M; ++n) {
      mbar_expect_tx(full_a + 8 * n, THREADS * 4);
      bulk_g2s(ring_a + n * THREADS * 4, fr + static_cast<int64_t>(1 + n) * HW, THREADS * 4, full_a + 8 * n);
    }
  }
  for (int i = 0; i < M; ++i) {
    const int s = i % DEPTH;
    const uint32_t par = (i / DEPTH) & 1;
    mbar_wait(full_a + 8 * s, par);
    const uint32_t w = *reinterpret_cast<const uint32_t*>(&ring[s][threadIdx.x * 4]);
    __syncwarp();
    if ((threadIdx.x & 31) == 0) mbar_arrive(empty_a + 8 * s);
    emit(vox, w);
    vox += HW;
    if (threadIdx.x == 0 && i + DEPTH < M) {      // refill this slot with frame i + DEPTH once every warp has released it
      mbar_wait(empty_a + 8 * s, par);
      mbar_expect_tx(full_a + 8 * s, THREADS * 4);
      bulk_g2s(ring_a + s * THREADS * 4, fr + static_cast<int64_t>(1 + i + DEPTH) * HW, THREADS * 4, full_a + 8 * s);
    }
  }
}

__global__ void copy_kernel(const float4* a, float4* b, int64_t n) {
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<int64_t>(gridDim.x) * blockDim.x) b[i] = a[i];
}

int main() {
  uint8_t* frames;
  float* voxel;
  const size_t fb = static_cast<size_t>(CLIPS) * NFRAMES * HW, vb = static_cast<size_t>(CLIPS) * (NFRAMES - 1) * HW * 4;
  cudaMalloc(&frames, fb);
  cudaMalloc(&voxel, vb);
  cudaMemset(frames, 7, fb);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  dim3 grid(HW / (THREADS * 4), CLIPS);
  const double bytes = static_cast<double>(CLIPS) * HW * (NFRAMES + (NFRAMES - 1) * 4.0);
  for (int which = 0; which < 3; ++which) {
    float best = 1e9f;
    for (int it = 0; it < 8; ++it) {
      cudaEventRecord(e0);
      if (which == 0) pattern_ldg<<<grid, THREADS>>>(frames, voxel);
      if (which == 1) pattern_bulk<<<grid, THREADS>>>(frames, voxel);
      if (which == 2) copy_kernel<<<148 * 8, 1024>>>(reinterpret_cast<const float4*>(voxel), reinterpret_cast<float4*>(voxel) + vb / 32, static_cast<int64_t>(vb / 32));
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      if (it >= 2 && ms < best) best = ms;
    }
    const double by = which == 2 ? static_cast<double>(vb) : bytes;      // copy: vb/2 read + vb/2 written
    printf("%-28s %7.3f ms  %7.0f GB/s  (%s)\n", which == 0 ? "pattern, LDG.32 ring" : which == 1 ? "pattern, cp.async.bulk ring" : "float4 copy", best,
           by / best / 1e6, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}

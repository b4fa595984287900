// Micro-experiment for the round-2 plan in DESIGN.md §8: can ONE TMA-loaded halo tile serve all 3x3 taps of an
// implicit-GEMM conv, by pointing the UMMA shared-memory descriptor at a window shifted by whole pixels?
//
// A halo tile of H x P pixels (P = row pitch in pixels, 64 FP16 channels = 128 B per pixel, SWIZZLE_128B) is loaded
// with one TMA box.  The A operand of an M=128 MMA is the window of 16 rows x 8 pixels whose top-left pixel is
// (dy, dx): row r = 8*y + x of the MMA is pixel (y + dy, x + dx), i.e. byte offset ((y+dy)*P + (x+dx))*128 ->
// descriptor start = base + (dy*P + dx)*128, stride-byte-offset (8-row group pitch) = P*128.  B is a 64x64 identity,
// so D = A and every output row names the pixel / 16-byte chunk the tensor core actually read.
//
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o halo_desc_probe halo_desc_probe.cu
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CHECK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

struct Cfg { int pitch, dy, dx, base_mode; };   // base_mode 0: descriptor base_offset = 0; 1: (start >> 7) & 7

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity))
    if (clock64() - t0 > 2000000000LL) __trap();          // ~1 s: never hang the box
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
               ::"r"(dst), "l"(tmap), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(dst), "l"(tmap), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}

__host__ __device__ constexpr uint32_t make_idesc(int bn) { return (1u << 4) | ((uint32_t)(bn >> 3) << 17) | ((uint32_t)(128 >> 4) << 24); }
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t sbo, uint32_t base_off) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFFu);
  d |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(base_off & 7u) << 49;
  d |= (uint64_t)2 << 61;
  return d;
}

constexpr int kRows = 18;                         // halo rows
constexpr int kA16 = kRows * 16 * 128;            // pitch-16 tile bytes
constexpr int kA10 = kRows * 10 * 128;            // pitch-10 tile bytes
constexpr int kA10Pad = (kA10 + 1023) / 1024 * 1024;

__global__ void __launch_bounds__(128, 1)
probe_kernel(const __grid_constant__ CUtensorMap map16, const __grid_constant__ CUtensorMap map10,
             const __grid_constant__ CUtensorMap mapB, const Cfg* cfgs, int ncfg, float* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA16 = smem;
  uint8_t* sA10 = smem + kA16;                    // 36864 = 36 KB: still 1024-aligned
  uint8_t* sB = sA10 + kA10Pad;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + 8192);
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 2);
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) {
    mbar_init(smem_u32(&bars[0]), 1);
    mbar_init(smem_u32(&bars[1]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"(64u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *s_tmem;
  if (tid == 0) {
    mbar_expect_tx(smem_u32(&bars[0]), kA16 + kA10 + 8192);
    tma_load_4d(smem_u32(sA16), &map16, smem_u32(&bars[0]), 0, 0, 0, 0);
    tma_load_4d(smem_u32(sA10), &map10, smem_u32(&bars[0]), 0, 0, 0, 0);
    tma_load_2d(smem_u32(sB), &mapB, smem_u32(&bars[0]), 0, 0);
  }
  mbar_wait(smem_u32(&bars[0]), 0);
  uint32_t phase = 0;
  for (int t = 0; t < ncfg; ++t) {
    const Cfg c = cfgs[t];
    if (tid == 0) {
      const uint32_t base = smem_u32(c.pitch == 16 ? sA16 : sA10);
      const uint32_t start = base + (uint32_t)(c.dy * c.pitch + c.dx) * 128u;
      const uint32_t boff = c.base_mode ? ((start >> 7) & 7u) : 0u;
      const uint64_t ad = make_desc(start, (uint32_t)c.pitch * 128u, boff);
      const uint64_t bd = make_desc(smem_u32(sB), 1024u, 0u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      for (int kk = 0; kk < 4; ++kk) umma_f16(tmem, ad + (uint64_t)(kk * 2), bd + (uint64_t)(kk * 2), make_idesc(64), kk ? 1u : 0u);
      umma_commit(smem_u32(&bars[1]));
    }
    mbar_wait(smem_u32(&bars[1]), phase);
    phase ^= 1;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t v0[32], v1[32];
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
    tmem_ld32(taddr, v0);
    tmem_ld32(taddr + 32, v1);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    float* o = out + ((size_t)t * 128 + tid) * 64;
    for (int j = 0; j < 32; ++j) { o[j] = __uint_as_float(v0[j]); o[32 + j] = __uint_as_float(v1[j]); }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
  }
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64u) : "memory");
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static float code(int q, int c) { return (float)(q - 144) + (float)(c >> 3) * 0.125f; }

int main() {
  void* fp = nullptr;
  cudaDriverEntryPointQueryResult qr;
  CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &qr));
  PFN_encodeTiled enc = (PFN_encodeTiled)fp;
  // halo tiles: pixel q = y*pitch + x, 64 channels
  auto make_tile = [&](int pitch, __half** d) {
    std::vector<__half> h((size_t)kRows * pitch * 64);
    for (int q = 0; q < kRows * pitch; ++q)
      for (int c = 0; c < 64; ++c) h[(size_t)q * 64 + c] = __float2half(code(q, c));
    CHECK(cudaMalloc(d, h.size() * 2));
    CHECK(cudaMemcpy(*d, h.data(), h.size() * 2, cudaMemcpyHostToDevice));
  };
  __half *d16, *d10, *dB;
  make_tile(16, &d16);
  make_tile(10, &d10);
  std::vector<__half> hb(64 * 64, __float2half(0.f));
  for (int i = 0; i < 64; ++i) hb[i * 64 + i] = __float2half(1.f);
  CHECK(cudaMalloc(&dB, hb.size() * 2));
  CHECK(cudaMemcpy(dB, hb.data(), hb.size() * 2, cudaMemcpyHostToDevice));
  CUtensorMap m16, m10, mB;
  auto enc_tile = [&](CUtensorMap* m, void* p, int pitch) {
    cuuint64_t dims[4] = {64, (cuuint64_t)pitch, (cuuint64_t)kRows, 1};
    cuuint64_t strides[3] = {128, (cuuint64_t)pitch * 128, (cuuint64_t)pitch * 128 * kRows};
    cuuint32_t box[4] = {64, (cuuint32_t)pitch, (cuuint32_t)kRows, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, p, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); exit(2); }
  };
  enc_tile(&m16, d16, 16);
  enc_tile(&m10, d10, 10);
  {
    cuuint64_t dims[2] = {64, 64};
    cuuint64_t strides[1] = {128};
    cuuint32_t box[2] = {64, 64};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&mB, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, dB, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode B failed %d\n", (int)r); exit(2); }
  }
  std::vector<Cfg> cfgs;
  for (int pitch : {16, 10})
    for (int bm : {0, 1})
      for (int dy = 0; dy < 3; ++dy)
        for (int dx = 0; dx < 3; ++dx) cfgs.push_back({pitch, dy, dx, bm});
  Cfg* dc;
  float* dout;
  CHECK(cudaMalloc(&dc, cfgs.size() * sizeof(Cfg)));
  CHECK(cudaMemcpy(dc, cfgs.data(), cfgs.size() * sizeof(Cfg), cudaMemcpyHostToDevice));
  CHECK(cudaMalloc(&dout, cfgs.size() * 128 * 64 * sizeof(float)));
  const int smem_bytes = 1024 + kA16 + kA10Pad + 8192 + 64;
  CHECK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
  probe_kernel<<<1, 128, smem_bytes>>>(m16, m10, mB, dc, (int)cfgs.size(), dout);
  CHECK(cudaDeviceSynchronize());
  std::vector<float> out(cfgs.size() * 128 * 64);
  CHECK(cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost));
  int n_ok = 0;
  for (size_t t = 0; t < cfgs.size(); ++t) {
    const Cfg c = cfgs[t];
    int bad = 0, first_r = -1, first_n = -1;
    for (int r = 0; r < 128; ++r) {
      const int y = r >> 3, x = r & 7, q = (y + c.dy) * c.pitch + x + c.dx;
      for (int n = 0; n < 64; ++n)
        if (out[(t * 128 + r) * 64 + n] != code(q, n)) { if (!bad) { first_r = r; first_n = n; } ++bad; }
    }
    printf("pitch %2d dy %d dx %d base_offset_mode %d : %s", c.pitch, c.dy, c.dx, c.base_mode, bad ? "MISMATCH" : "ok");
    if (bad) {
      printf("  (%d of 8192 wrong; first at row %d col %d)\n    row: got pixel/chunk for cols 0,8,16,.. :", bad, first_r, first_n);
      for (int r = first_r; r < first_r + 3 && r < 128; ++r) {
        printf("\n    r=%3d want q=%3d :", r, ((r >> 3) + c.dy) * c.pitch + (r & 7) + c.dx);
        for (int n = 0; n < 64; n += 8) {
          const float v = out[(t * 128 + r) * 64 + n];
          const float fl = floorf(v);
          printf(" q%d/c%d", (int)fl + 144, (int)((v - fl) * 8.f));
        }
      }
    } else ++n_ok;
    printf("\n");
  }
  printf("%d of %zu configurations exact\n", n_ok, cfgs.size());
  return 0;
}

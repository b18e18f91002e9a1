// K1 x pass with TMA-staged tiles (opt-in: b2t_edt_config_xpass(1) / B2T_EDT_XTMA=1).
//
// A CTA of eight warps takes a tile of eight rows: ONE elected thread issues cp.async.bulk.tensor.2d loads of the label
// tile (boxes of 256 labels x 8 rows, the hardware's limit per box dimension) into shared memory behind an mbarrier,
// every warp then runs the row algorithm of edt_pass_x_v2_kernel on its row out of shared memory (16 labels per lane in
// registers, run boundaries by clz / ffs and two warp scans), writes its 16 results back to shared memory, and one
// thread sends the result tile to global memory with cp.async.bulk.tensor.2d stores.  No per-lane global address
// arithmetic, no LDG / STG: the SASS of this kernel has UTMALDG / UTMASTG.  Same arithmetic, same bits as the v2 kernel.
#pragma once
#include <cuda.h>

namespace xtma {

constexpr int kRows = 8;          // rows per tile = warps per CTA
constexpr int kBoxW = 256;        // labels per box row (box dimensions are limited to 256 elements)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "XTMA_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra XTMA_DONE;\n"
      "bra XTMA_WAIT;\n"
      "XTMA_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, unsigned long long* bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                   smem_u32(dst)),
               "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map), "r"(smem_u32(src)),
               "r"(c0), "r"(c1)
               : "memory");
}

// one warp, one row: labels from the tile in shared memory (NB boxes of 256), results to the result tile
template <int NB>
__device__ __forceinline__ void row_from_tile(const uint32_t (*tin)[kRows][kBoxW], float (*tout)[kRows][kBoxW], int wrp, int lane,
                                              int sx, float w, int black_border) {
  constexpr int SEG = NB * kBoxW / 32;
  const int p0 = lane * SEG;
  const int bx = p0 / kBoxW, px = p0 % kBoxW;           // a lane's 16 labels lie inside one box
  const uint32_t* lrow = &tin[bx][wrp][px];
  uint32_t lab[SEG];
#pragma unroll
  for (int q = 0; q < SEG / 4; q++) {
    const uint4 v = *reinterpret_cast<const uint4*>(lrow + 4 * q);
    lab[4 * q] = v.x; lab[4 * q + 1] = v.y; lab[4 * q + 2] = v.z; lab[4 * q + 3] = v.w;
  }
  uint32_t prev = __shfl_up_sync(0xffffffffu, lab[SEG - 1], 1);
  uint32_t brk = 0;
#pragma unroll
  for (int j = 0; j < SEG; j++) {
    const int p = p0 + j;
    const uint32_t pl = (j == 0) ? prev : lab[j - 1];
    if (p == 0 || lab[j] != pl) brk |= 1u << j;
  }
  int last_in = brk ? p0 + 31 - __clz(brk) : -1;
  int first_in = brk ? p0 + __ffs(brk) - 1 : sx;
  int lastb = last_in, firstb = first_in;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int a = __shfl_up_sync(0xffffffffu, lastb, o);
    const int b = __shfl_down_sync(0xffffffffu, firstb, o);
    if (lane >= o) lastb = max(lastb, a);
    if (lane + o < 32) firstb = min(firstb, b);
  }
  int left_in = __shfl_up_sync(0xffffffffu, lastb, 1);
  int right_in = __shfl_down_sync(0xffffffffu, firstb, 1);
  if (lane == 0) left_in = 0;
  if (lane == 31) right_in = sx;
  float val[SEG];
#pragma unroll
  for (int j = 0; j < SEG; j++) {
    const int p = p0 + j;
    const uint32_t le = brk & (0xffffffffu >> (31 - j));
    const uint32_t gt = (j == 31) ? 0u : (brk & (0xffffffffu << (j + 1)));
    const int s = le ? (p0 + 31 - __clz(le)) : left_in;
    const int nx = gt ? (p0 + __ffs(gt) - 1) : right_in;
    float v = 0.0f;
    if (lab[j] != 0) {
      const bool lok = (s > 0) || black_border, rok = (nx < sx) || black_border;
      if (lok || rok) {
        const int dl = p - s + 1, dr = nx - p;
        const int d = lok ? (rok ? min(dl, dr) : dl) : dr;
        const float fd = __fmul_rn((float)d, w);
        v = __fmul_rn(fd, fd);
      } else {
        v = __int_as_float(0x7f800000);
      }
    }
    val[j] = v;
  }
  float* orow = &tout[bx][wrp][px];
#pragma unroll
  for (int q = 0; q < SEG / 4; q++)
    *reinterpret_cast<float4*>(orow + 4 * q) = make_float4(val[4 * q], val[4 * q + 1], val[4 * q + 2], val[4 * q + 3]);
}

// NB = boxes per row (sx = NB * 256), SEG = labels per lane (sx = 32 * SEG)
template <int NB>
__global__ void __launch_bounds__(kRows * 32) edt_pass_x_tma_kernel(const __grid_constant__ CUtensorMap tm_in,
                                                                     const __grid_constant__ CUtensorMap tm_out, int sx,
                                                                     int64_t nrows, float w, int black_border) {
  __shared__ __align__(128) uint32_t s_in[NB][kRows][kBoxW];
  __shared__ __align__(128) float s_out[NB][kRows][kBoxW];
  __shared__ __align__(8) unsigned long long bar;
  const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
  const int64_t row0 = (int64_t)blockIdx.x * kRows;
  if (threadIdx.x == 0) mbar_init(&bar, 1);
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_expect_tx(&bar, (uint32_t)(NB * kRows * kBoxW * sizeof(uint32_t)));
#pragma unroll
    for (int b = 0; b < NB; b++) tma_load_2d(&s_in[b][0][0], &tm_in, &bar, b * kBoxW, (int)row0);
  }
  mbar_wait(&bar, 0);
  row_from_tile<NB>(s_in, s_out, wrp, lane, sx, w, black_border);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");       // generic-proxy writes -> visible to the bulk store
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int b = 0; b < NB; b++) tma_store_2d(&tm_out, &s_out[b][0][0], b * kBoxW, (int)row0);
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");    // shared memory stays alive until it has been read
  }
}

// The same pass as a PERSISTENT kernel with a two-stage ring: while the warps work on tile i out of stage i & 1, the
// loads of tile i + 1 are in flight into the other stage and the store of tile i - 1 drains; two block barriers per tile.
template <int NB>
__global__ void __launch_bounds__(kRows * 32) edt_pass_x_tma2_kernel(const __grid_constant__ CUtensorMap tm_in,
                                                                      const __grid_constant__ CUtensorMap tm_out, int sx,
                                                                      int64_t nrows, float w, int black_border) {
  extern __shared__ __align__(128) unsigned char xtma_smem[];
  typedef uint32_t TileIn[NB][kRows][kBoxW];
  typedef float TileOut[NB][kRows][kBoxW];
  TileIn* s_in = reinterpret_cast<TileIn*>(xtma_smem);                              // [2]
  TileOut* s_out = reinterpret_cast<TileOut*>(xtma_smem + 2 * sizeof(TileIn));      // [2]
  __shared__ __align__(8) unsigned long long full[2];
  const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
  const int64_t ntiles = (nrows + kRows - 1) / kRows;
  constexpr uint32_t kBytes = (uint32_t)(NB * kRows * kBoxW * sizeof(uint32_t));
  if (threadIdx.x == 0) { mbar_init(&full[0], 1); mbar_init(&full[1], 1); }
  __syncthreads();
  int64_t t = blockIdx.x;
  if (t >= ntiles) return;
  if (threadIdx.x == 0) {
    mbar_expect_tx(&full[0], kBytes);
#pragma unroll
    for (int b = 0; b < NB; b++) tma_load_2d(&s_in[0][b][0][0], &tm_in, &full[0], b * kBoxW, (int)(t * kRows));
  }
  for (uint32_t it = 0; t < ntiles; it++, t += gridDim.x) {
    const uint32_t st = it & 1u;
    if (threadIdx.x == 0) {
      asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");    // the store that read s_out[st] two tiles ago is done with it
      const int64_t tn = t + gridDim.x;
      if (tn < ntiles) {                                                // s_in[st ^ 1] was read in the last iteration: free
        mbar_expect_tx(&full[st ^ 1u], kBytes);
#pragma unroll
        for (int b = 0; b < NB; b++) tma_load_2d(&s_in[st ^ 1u][b][0][0], &tm_in, &full[st ^ 1u], b * kBoxW, (int)(tn * kRows));
      }
    }
    __syncthreads();
    mbar_wait(&full[st], (it >> 1) & 1u);
    row_from_tile<NB>(s_in[st], s_out[st], wrp, lane, sx, w, black_border);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
      for (int b = 0; b < NB; b++) tma_store_2d(&tm_out, &s_out[st][b][0][0], b * kBoxW, (int)(t * kRows));
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
  }
  if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeFn encoder() {
  static EncodeFn fn = []() -> EncodeFn {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
      return nullptr;
    return reinterpret_cast<EncodeFn>(p);
  }();
  return fn;
}

// rows x sx array of 4-byte elements, boxes of 256 x 8
inline bool make_map(CUtensorMap* m, CUtensorMapDataType dt, const void* base, int64_t sx, int64_t nrows) {
  EncodeFn f = encoder();
  if (!f) return false;
  const cuuint64_t dims[2] = {(cuuint64_t)sx, (cuuint64_t)nrows};
  const cuuint64_t strides[1] = {(cuuint64_t)sx * 4};
  const cuuint32_t box[2] = {(cuuint32_t)kBoxW, (cuuint32_t)kRows};
  const cuuint32_t es[2] = {1, 1};
  return f(m, dt, 2, const_cast<void*>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
           CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

inline unsigned long long& launches() { static unsigned long long n = 0; return n; }

// returns false when this form does not apply (the caller launches the v2 kernel instead)
inline bool launch(const uint32_t* labels, float* out, int64_t sx, int64_t nrows, float w, int black_border, cudaStream_t st,
                   int form = 1) {
  if (sx != 256 && sx != 512) return false;
  CUtensorMap tin, tout;
  if (!make_map(&tin, CU_TENSOR_MAP_DATA_TYPE_UINT32, labels, sx, nrows)) return false;
  if (!make_map(&tout, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, out, sx, nrows)) return false;
  if (form == 2) {                                          // persistent, two-stage ring (dynamic shared memory)
    const size_t smem1 = 4 * (size_t)1 * kRows * kBoxW * 4, smem2 = 4 * (size_t)2 * kRows * kBoxW * 4;
    static bool attr = false;
    if (!attr) {
      if (cudaFuncSetAttribute(edt_pass_x_tma2_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1) != cudaSuccess ||
          cudaFuncSetAttribute(edt_pass_x_tma2_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2) != cudaSuccess)
        return false;
      attr = true;
    }
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int64_t ntiles = (nrows + kRows - 1) / kRows;
    const int per_sm = sx == 256 ? 6 : 3;
    const unsigned grid = (unsigned)(ntiles < (int64_t)sms * per_sm ? ntiles : (int64_t)sms * per_sm);
    if (sx == 256) edt_pass_x_tma2_kernel<1><<<grid, kRows * 32, smem1, st>>>(tin, tout, (int)sx, nrows, w, black_border);
    else edt_pass_x_tma2_kernel<2><<<grid, kRows * 32, smem2, st>>>(tin, tout, (int)sx, nrows, w, black_border);
    launches()++;
    return true;
  }
  const unsigned blocks = (unsigned)((nrows + kRows - 1) / kRows);
  if (sx == 256) edt_pass_x_tma_kernel<1><<<blocks, kRows * 32, 0, st>>>(tin, tout, (int)sx, nrows, w, black_border);
  else edt_pass_x_tma_kernel<2><<<blocks, kRows * 32, 0, st>>>(tin, tout, (int)sx, nrows, w, black_border);
  launches()++;
  return true;
}

}  // namespace xtma

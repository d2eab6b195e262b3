// Tensor-core GEMM for sm_100a: C[M,N] (fp32) = A[M,K] B[N,K]^T (+bias) (ReLU), tcgen05.mma with the accumulator in
// TMEM, operands staged by TMA (cp.async.bulk.tensor, 128-byte swizzle), mbarrier pipelines, persistent CTAs.
//
// fp32-accurate mode ("bf16x3"): every fp32 operand is pre-split (caphn_split_bf16*) into bf16 hi = rn(x) and
// lo = rn(x - hi); the kernel issues three MMAs per k-slice, hi*hi + hi*lo + lo*hi, accumulating in fp32.  The dropped
// lo*lo term and the residual of the split are ~2^-17 relative, i.e. ~1e-5 on the result -- inside the 1e-4 parity
// budget of BASELINE.json (plain TF32 would be ~1e-3).  With lo == NULL the same kernel runs a plain bf16 GEMM.
//
// Replaces the large addmm calls behind nn.Linear in the reference: vocabulary projection self.fc / fc_out
// (models/decoderlstm.py:105, later.py:442) forward and both backward products, feature_fc (models/decoderlstm.py:61).
//
// Warp roles (256 threads, 1 CTA/SM): warp 0 = TMA producer, warp 1 = MMA issuer (one elected lane), warp 2 = TMEM
// allocator, warps 4-7 = epilogue (tcgen05.ld -> smem transpose -> coalesced global stores).  Two TMEM accumulator
// stages (2 x 128 columns) let the MMAs of tile i+1 overlap the epilogue of tile i.
#include "common.cuh"
#include <cuda.h>
#include <cuda_bf16.h>

namespace caphn {
namespace tc {

constexpr int BM = 128, BN = 128, BK = 64;
constexpr int TILE_BYTES = BM * BK * 2;  // 16 KiB: one 128 x 64 bf16 operand tile (128-byte rows, SWIZZLE_128B)
constexpr int THREADS = 256;
constexpr int STG_FLOATS = 4 * 32 * 33;
constexpr uint32_t TMEM_COLS = 256;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// Bounded wait: a protocol bug traps (kernel error) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t done = 0;
    const long long t0 = clock64();
    while (true) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
        if (done) break;
        if (clock64() - t0 > 4000000000LL) __trap();
    }
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* tmap, uint64_t* bar, int x, int y) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(x), "r"(y)
        : "memory");
}
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor layout): start address >> 4 in
// bits [0,14), LBO = 1 (16-byte units; fixed for swizzled K-major), SBO = 1024 B (8 rows x 128 B) in bits [32,46),
// version = 1 (Blackwell) in bits [46,48), layout type 2 (SWIZZLE_128B) in bits [61,64).
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// Instruction descriptor for kind::f16: D = F32 (bits [4,6) = 1), A = B = BF16 (bits [7,10) = [10,13) = 1), both K-major,
// N >> 3 in bits [17,23), M >> 4 in bits [24,29).
__device__ __forceinline__ uint32_t make_idesc(int m, int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

template <bool SPLIT>
struct Cfg {
    static constexpr int NT = SPLIT ? 4 : 2;          // operand tiles per pipeline stage
    static constexpr int STAGES = SPLIT ? 3 : 6;
    static constexpr int STAGE_BYTES = NT * TILE_BYTES;
    static constexpr size_t SMEM = (size_t)STAGES * STAGE_BYTES + STG_FLOATS * 4 + (2 * STAGES + 4) * 8 + 16 + 1024;
};

template <bool SPLIT>
__global__ void __launch_bounds__(THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmAh, const __grid_constant__ CUtensorMap tmAl,
               const __grid_constant__ CUtensorMap tmBh, const __grid_constant__ CUtensorMap tmBl,
               float* __restrict__ C, long ldc, const float* __restrict__ bias, int M, int N, int num_kb, int relu) {
    using G = Cfg<SPLIT>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* tiles = smem;
    float* stg = reinterpret_cast<float*>(smem + (size_t)G::STAGES * G::STAGE_BYTES);
    uint64_t* full = reinterpret_cast<uint64_t*>(stg + STG_FLOATS);
    uint64_t* empty = full + G::STAGES;
    uint64_t* tfull = empty + G::STAGES;
    uint64_t* tempty = tfull + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tiles_m = (M + BM - 1) / BM, tiles_n = (N + BN - 1) / BN;
    const int num_tiles = tiles_m * tiles_n;

    if (threadIdx.x == 0) {
        for (int s = 0; s < G::STAGES; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(tfull + s, 1); mbar_init(tempty + s, 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"(TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const int mb = tile % tiles_m, nb = tile / tiles_m;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(empty + stage, phase ^ 1);
                    mbar_expect_tx(full + stage, G::STAGE_BYTES);
                    uint8_t* st = tiles + (size_t)stage * G::STAGE_BYTES;
                    tma_load_2d(st, &tmAh, full + stage, kb * BK, mb * BM);
                    tma_load_2d(st + TILE_BYTES, &tmBh, full + stage, kb * BK, nb * BN);
                    if (SPLIT) {
                        tma_load_2d(st + 2 * TILE_BYTES, &tmAl, full + stage, kb * BK, mb * BM);
                        tma_load_2d(st + 3 * TILE_BYTES, &tmBl, full + stage, kb * BK, nb * BN);
                    }
                    if (++stage == G::STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = make_idesc(BM, BN);
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
                const int as = it & 1;
                const uint32_t aph = (it >> 1) & 1;
                mbar_wait(tempty + as, aph ^ 1);
                tcgen05_fence_after();
                const uint32_t tmem_d = tmem_base + (uint32_t)(as * BN);
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(full + stage, phase);
                    tcgen05_fence_after();
                    const uint32_t sbase = smem_u32(tiles + (size_t)stage * G::STAGE_BYTES);
                    const uint64_t dAh = make_desc(sbase), dBh = make_desc(sbase + TILE_BYTES);
                    const uint64_t dAl = make_desc(sbase + 2 * TILE_BYTES), dBl = make_desc(sbase + 3 * TILE_BYTES);
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k) {
                        const uint64_t ko = (uint64_t)((k * 16 * 2) >> 4);  // advance 32 bytes along K inside the swizzle atom
                        umma_f16(tmem_d, dAh + ko, dBh + ko, idesc, (kb | k) ? 1u : 0u);
                        if (SPLIT) {
                            umma_f16(tmem_d, dAh + ko, dBl + ko, idesc, 1u);
                            umma_f16(tmem_d, dAl + ko, dBh + ko, idesc, 1u);
                        }
                    }
                    umma_commit(empty + stage);   // frees this smem stage once the MMAs above have read it
                    if (++stage == G::STAGES) { stage = 0; phase ^= 1; }
                }
                umma_commit(tfull + as);          // accumulator complete -> epilogue
            }
        }
    } else if (warp >= 4) {
        const int q = warp - 4;                   // TMEM lane quadrant this warp may access
        float* st = stg + q * (32 * 33);
        int it = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
            const int mb = tile % tiles_m, nb = tile / tiles_m;
            const int as = it & 1;
            const uint32_t aph = (it >> 1) & 1;
            mbar_wait(tfull + as, aph);
            tcgen05_fence_after();
            const int row0 = mb * BM + q * 32;
#pragma unroll 1
            for (int c = 0; c < BN / 32; ++c) {
                uint32_t r[32];
                tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * BN + c * 32), r);
                tmem_ld_wait();
                if (c == BN / 32 - 1) {
                    tcgen05_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(tempty + as);   // TMEM stage drained (4 warps -> count 4)
                }
#pragma unroll
                for (int j = 0; j < 32; ++j) st[lane * 33 + j] = __uint_as_float(r[j]);
                __syncwarp();
                const int col = nb * BN + c * 32 + lane;
                const bool colok = col < N;
                const float bv = (bias != nullptr && colok) ? bias[col] : 0.f;
#pragma unroll 8
                for (int rr = 0; rr < 32; ++rr) {
                    const int row = row0 + rr;
                    if (row < M && colok) {
                        float v = st[rr * 33 + lane] + bv;
                        if (relu) v = fmaxf(v, 0.f);
                        C[(long)row * ldc + col] = v;
                    }
                }
                __syncwarp();
            }
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 2) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// hi[r, c] = bf16(src[r, c]);  lo[r, c] = bf16(src - hi);  columns [C, Kp) are zeroed.   (lo may be null)
__global__ void split_bf16_kernel(const float* __restrict__ src, long lds, long R, int C, __nv_bfloat16* __restrict__ hi,
                                  __nv_bfloat16* __restrict__ lo, long Kp) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;  // one thread per pair of columns
    const long half = Kp >> 1;
    if (i >= R * half) return;
    const long r = i / half;
    const int c = (int)(i - r * half) * 2;
    float x0 = 0.f, x1 = 0.f;
    if (c < C) x0 = src[r * lds + c];
    if (c + 1 < C) x1 = src[r * lds + c + 1];
    const __nv_bfloat16 h0 = __float2bfloat16_rn(x0), h1 = __float2bfloat16_rn(x1);
    __nv_bfloat162 hv; hv.x = h0; hv.y = h1;
    *reinterpret_cast<__nv_bfloat162*>(hi + r * Kp + c) = hv;
    if (lo) {
        __nv_bfloat162 lv;
        lv.x = __float2bfloat16_rn(x0 - __bfloat162float(h0));
        lv.y = __float2bfloat16_rn(x1 - __bfloat162float(h1));
        *reinterpret_cast<__nv_bfloat162*>(lo + r * Kp + c) = lv;
    }
}

// Transposing split: hi/lo [C, Rp] with hi[c, r] = bf16(src[r, c]); columns [R, Rp) zeroed.
__global__ void split_bf16_t_kernel(const float* __restrict__ src, long lds, int R, int C, __nv_bfloat16* __restrict__ hi,
                                    __nv_bfloat16* __restrict__ lo, long Rp) {
    __shared__ float tile[32][33];
    const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int r = r0 + i, c = c0 + threadIdx.x;
        tile[i][threadIdx.x] = (r < R && c < C) ? src[(long)r * lds + c] : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int c = c0 + i, r = r0 + threadIdx.x;
        if (c < C && r < Rp) {
            const float x = tile[threadIdx.x][i];
            const __nv_bfloat16 h = __float2bfloat16_rn(x);
            hi[(long)c * Rp + r] = h;
            if (lo) lo[(long)c * Rp + r] = __float2bfloat16_rn(x - __bfloat162float(h));
        }
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// 2-D map over a row-major bf16 matrix [rows, Kp]: box = 64 (K) x 128 (rows), 128-byte swizzle, zero fill out of bounds.
static int make_map(CUtensorMap* m, const void* base, long rows, long Kp) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return CAPHN_EINVAL;
    cuuint64_t dims[2] = {(cuuint64_t)Kp, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)Kp * 2};
    cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)BM};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? CAPHN_OK : 1000 + (int)r;
}

}  // namespace tc
}  // namespace caphn

using namespace caphn;

extern "C" {

// hi/lo [R, Kp] bf16 (Kp % 64 == 0, Kp >= C): the bf16x3 operand format of caphn_gemm_tc.  lo may be NULL (bf16 mode).
int caphn_split_bf16(const float* src, long lds, long R, int C, void* hi, void* lo, long Kp, void* stream) {
    if (R <= 0 || C <= 0 || Kp < C || (Kp & 63)) return CAPHN_EINVAL;
    const long n = R * (Kp >> 1);
    tc::split_bf16_kernel<<<ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(
        src, lds, R, C, (__nv_bfloat16*)hi, (__nv_bfloat16*)lo, Kp);
    CAPHN_RETURN_LAST();
}

// Transposing variant: src [R, C] fp32 -> hi/lo [C, Rp] bf16 (Rp % 64 == 0, Rp >= R).
int caphn_split_bf16_t(const float* src, long lds, int R, int C, void* hi, void* lo, long Rp, void* stream) {
    if (R <= 0 || C <= 0 || Rp < R || (Rp & 63)) return CAPHN_EINVAL;
    dim3 grid(ceil_div(C, 32), ceil_div(Rp, 32));
    tc::split_bf16_t_kernel<<<grid, dim3(32, 8), 0, (cudaStream_t)stream>>>(src, lds, R, C, (__nv_bfloat16*)hi,
                                                                            (__nv_bfloat16*)lo, Rp);
    CAPHN_RETURN_LAST();
}

// C[M,N] (fp32, row stride ldc) = A B^T (+bias[n]) (ReLU) on the tensor cores.  A = (Ahi, Alo) [M, Kp], B = (Bhi, Blo)
// [N, Kp] in the split format above (16-byte aligned, Kp % 64 == 0).  Alo == Blo == NULL selects plain bf16.
int caphn_gemm_tc(const void* Ahi, const void* Alo, const void* Bhi, const void* Blo, long Kp, float* C, long ldc,
                  const float* bias, int M, int N, int relu, void* stream) {
    if (M <= 0 || N <= 0 || Kp <= 0 || (Kp & 63) || ((Alo == nullptr) != (Blo == nullptr))) return CAPHN_EINVAL;
    if (((uintptr_t)Ahi & 15) || ((uintptr_t)Bhi & 15) || ((uintptr_t)Alo & 15) || ((uintptr_t)Blo & 15))
        return CAPHN_EINVAL;
    const bool split = Alo != nullptr;
    CUtensorMap mAh, mAl, mBh, mBl;
    int rc;
    if ((rc = tc::make_map(&mAh, Ahi, M, Kp))) return rc;
    if ((rc = tc::make_map(&mBh, Bhi, N, Kp))) return rc;
    if ((rc = tc::make_map(&mAl, split ? Alo : Ahi, M, Kp))) return rc;
    if ((rc = tc::make_map(&mBl, split ? Blo : Bhi, N, Kp))) return rc;
    const int tiles = ceil_div(M, tc::BM) * ceil_div(N, tc::BN);
    const int grid = tiles < kNumSMs ? tiles : kNumSMs;
    const int num_kb = (int)(Kp / tc::BK);
    cudaStream_t st = (cudaStream_t)stream;
    if (split) {
        CAPHN_CHECK(cudaFuncSetAttribute(tc::gemm_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)tc::Cfg<true>::SMEM));
        tc::gemm_tc_kernel<true><<<grid, tc::THREADS, tc::Cfg<true>::SMEM, st>>>(mAh, mAl, mBh, mBl, C, ldc, bias, M, N,
                                                                                 num_kb, relu);
    } else {
        CAPHN_CHECK(cudaFuncSetAttribute(tc::gemm_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)tc::Cfg<false>::SMEM));
        tc::gemm_tc_kernel<false><<<grid, tc::THREADS, tc::Cfg<false>::SMEM, st>>>(mAh, mAl, mBh, mBl, C, ldc, bias, M, N,
                                                                                  num_kb, relu);
    }
    CAPHN_RETURN_LAST();
}

}  // extern "C"

// Shared pieces of the step-split recurrence kernels (attgru_step.cu forward, attgru_step_bwd.cu backward): async-copy
// helpers (mbarrier + 1-D bulk TMA, 16-byte cp.async), programmatic dependent launch, the warp-MMA inner loop over
// bf16 hi/lo operand rows, and the PDL launch wrapper.
#pragma once
#include "seq_common.cuh"
#include "mma_common.cuh"

namespace caphn {

constexpr int ST_MAXKT = 13;     // register-resident A fragments: K <= 208 (26 x uint4 per lane)

// ------------------------------------------------------------------------------------------------------------------
// async-copy helpers (mbarrier + 1-D bulk TMA for the K tile of X, 16-byte cp.async for the operand rows of Y)
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t st_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void st_mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(st_smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void st_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(st_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void st_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(st_smem_u32(dst)), "l"(src), "r"(bytes), "r"(st_smem_u32(bar)) : "memory");
}
// bounded wait: a protocol bug traps instead of hanging the GPU
__device__ __forceinline__ void st_mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = st_smem_u32(bar);
    uint32_t done = 0;
    const long long t0 = clock64();
    while (true) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(addr), "r"(parity) : "memory");
        if (done) break;
        if (clock64() - t0 > 2000000000LL) __trap();
    }
}
__device__ __forceinline__ uint4 st_ldg_u4(const uint4* p) {   // volatile: keeps the prefetch distance the source states
    uint4 r;
    asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ void st_cp_async16(void* dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(st_smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void st_cp_async_wait_all() {
    asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}

__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// B-operand rows of one plane pair (hi, lo) -> shared memory [2][NB][KP], 16-byte async copies; rows >= valid are zeroed
template <int NB, int THREADS>
__device__ __forceinline__ void stage_rows_async(const __nv_bfloat16* __restrict__ src, long plane, long row0, int valid,
                                                 int KP, __nv_bfloat16* dst, int tid) {
    const int CPR = KP >> 3;                       // 16-byte chunks per row (KP = 16 n + 8)
    for (int i = tid; i < 2 * NB * CPR; i += THREADS) {
        const int pl = i / (NB * CPR), r = i - pl * (NB * CPR);
        const int row = r / CPR, ch = r - row * CPR;
        __nv_bfloat16* d = dst + ((long)pl * NB + row) * KP + ch * 8;
        if (row < valid) st_cp_async16(d, src + pl * plane + (row0 + row) * KP + ch * 8);
        else *reinterpret_cast<uint4*>(d) = make_uint4(0u, 0u, 0u, 0u);
    }
}

// B-operand rows [row0, row0+valid) of `nplanes` planes -> shared memory [nplanes][NB][KP] by one bulk TMA copy per plane
// (the rows of a plane are contiguous); rows >= valid are zero-filled by the threads.  Thread 0 issues; everybody must
// then wait on `bar` (phase 0) before reading.
template <int NB, int THREADS>
__device__ __forceinline__ void stage_rows_bulk(const __nv_bfloat16* __restrict__ src, long plane, long row0, int valid,
                                                int KP, int nplanes, __nv_bfloat16* dst, uint64_t* bar, int tid) {
    if (tid == 0) {
        const uint32_t bytes = (uint32_t)valid * KP * 2;
        st_mbar_expect_tx(bar, bytes * nplanes);
        for (int pl = 0; pl < nplanes; ++pl) st_bulk_g2s(dst + (long)pl * NB * KP, src + pl * plane + row0 * KP, bytes, bar);
    }
    if (valid < NB) {
        const int CPR = KP >> 3;
        for (int i = tid; i < nplanes * (NB - valid) * CPR; i += THREADS) {
            const int pl = i / ((NB - valid) * CPR), r = i - pl * ((NB - valid) * CPR);
            *reinterpret_cast<uint4*>(dst + ((long)pl * NB + valid) * KP + r * 8) = make_uint4(0u, 0u, 0u, 0u);
        }
    }
}

// acc[nt] += A(hi,lo fragments of 13 k-tiles) . B(rows nt*8.. of the staged operand)
template <int NT>
__device__ __forceinline__ void warp_mma_rows(const uint4 (&ah)[ST_MAXKT], const uint4 (&al)[ST_MAXKT], int NKT,
                                              const __nv_bfloat16* bhi, const __nv_bfloat16* blo, int KP, int lane,
                                              float (&acc)[NT][4]) {
    const int nrow = lane >> 2, kc = (lane & 3) * 2;
#pragma unroll
    for (int kt = 0; kt < ST_MAXKT; ++kt) {
        if (kt < NKT) {
            const uint32_t fh[4] = {ah[kt].x, ah[kt].y, ah[kt].z, ah[kt].w};
            const uint32_t fl[4] = {al[kt].x, al[kt].y, al[kt].z, al[kt].w};
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
                const uint32_t* ph = reinterpret_cast<const uint32_t*>(bhi + (nt * 8 + nrow) * KP + kt * 16 + kc);
                const uint32_t* pl = reinterpret_cast<const uint32_t*>(blo + (nt * 8 + nrow) * KP + kt * 16 + kc);
                const uint32_t h0 = ph[0], h1 = ph[4], l0 = pl[0], l1 = pl[4];
                mma_bf16(acc[nt], fh, h0, h1);
                mma_bf16(acc[nt], fh, l0, l1);
                mma_bf16(acc[nt], fl, h0, h1);
            }
        }
    }
}

static inline size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

template <typename Kern, typename Arg>
static cudaError_t launch_pdl(Kern kern, dim3 grid, dim3 block, size_t smem, cudaStream_t stream, bool pdl, const Arg& arg) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, arg);
}

}  // namespace caphn

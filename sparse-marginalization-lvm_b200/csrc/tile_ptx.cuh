// PTX wrappers shared by the TMA-staged tile kernels (sparsemax_tile.cu, compact_tile.cu):
// mbarrier, 1-D bulk async copies (cp.async.bulk = TMA without a tensor map) and
// shared-memory accesses by 32-bit shared address.
#pragma once

#include <stdint.h>

namespace smlvm {

constexpr int kTileRows = 32;  // rows per tile: one row per lane of the owning warp

// ---- PTX: mbarrier + bulk async copies -----------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p)
{
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!ok);
}
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
        "l"(src), "r"(bytes), "r"(bar)
        : "memory");
}
__device__ __forceinline__ void bulk_store(void* dst, uint32_t src, uint32_t bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }


// shared-memory accesses by 32-bit shared address (volatile: the row is re-read and overwritten)
__device__ __forceinline__ float4 lds128(uint32_t a)
{
    float4 r;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "r"(a));
    return r;
}
__device__ __forceinline__ float lds32(uint32_t a)
{
    float r;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(r) : "r"(a));
    return r;
}
__device__ __forceinline__ void sts128(uint32_t a, float4 v)
{
    asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void sts32(uint32_t a, float v)
{
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory");
}

// bit (4*i + e) of the swizzled mask belongs to column (4*i + e) ^ (lsw << 2): permute the
// nibbles of the K-bit mask by XOR-ing their index with lsw (one butterfly level per bit).
template <int MW>
__device__ __forceinline__ void unswizzle_mask(uint32_t (&m)[MW], uint32_t lsw)
{
#pragma unroll
    for (int w = 0; w < MW; ++w) {
        uint32_t v = m[w];
        uint32_t t = ((v & 0xF0F0F0F0u) >> 4) | ((v & 0x0F0F0F0Fu) << 4);
        v = (lsw & 1u) ? t : v;
        t = __byte_perm(v, 0, 0x2301);
        v = (lsw & 2u) ? t : v;
        t = __byte_perm(v, 0, 0x1032);
        v = (lsw & 4u) ? t : v;
        m[w] = v;
    }
    if (MW >= 2) {
#pragma unroll
        for (int w = 0; w < MW; w += 2) {
            const uint32_t a = m[w], b = m[w + 1 < MW ? w + 1 : w];
            const bool sw = (lsw & 8u) != 0;
            m[w] = sw ? b : a;
            if (w + 1 < MW) m[w + 1] = sw ? a : b;
        }
    }
    if (MW >= 4) {
#pragma unroll
        for (int w = 0; w < MW; ++w) {
            if ((w & 2) == 0 && w + 2 < MW) {
                const uint32_t a = m[w], b = m[w + 2];
                const bool sw = (lsw & 16u) != 0;
                m[w] = sw ? b : a;
                m[w + 2] = sw ? a : b;
            }
        }
    }
}


}  // namespace smlvm

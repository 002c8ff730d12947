// umma.cuh -- tcgen05 / TMEM / bulk-copy helpers shared by the tensor-core kernels (scorer_head.cu, feature_tower.cu):
// shared-memory matrix descriptors (K-major, no swizzle: [K/8][rows][16 bytes] core-matrix layout), the kind::f16 instruction
// descriptor, MMA issue / commit, TMEM loads, and the bf16 hi/lo split that gives fp32-grade products from bf16 MMAs.
// Field layouts follow the PTX ISA's tcgen05 matrix / instruction descriptors (as encoded by cute/arch/mma_sm100_desc.hpp).
#pragma once

#include "common.cuh"
#include "tma.cuh"

constexpr int SH_M = 128;                    // rows per tile (MMA M)
constexpr int SH_ACHUNK = SH_M * 16;         // one 8-wide K chunk of an A operand: [128 rows][8 bf16]
constexpr unsigned SH_SPIN_LIMIT = 1u << 28; // a wait that long is a protocol bug: trap instead of hanging the GPU

__device__ __forceinline__ void sh_wait(uint64_t *bar, uint32_t parity)
{
	unsigned spins = 0;
	while (!mbar_try_wait(bar, parity))
		if (++spins > SH_SPIN_LIMIT) __trap();
}
__device__ __forceinline__ void sh_arrive(uint64_t *bar)
{
	asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tma_smem_addr(bar)) : "memory");
}
// shared-memory matrix descriptor, K-major, no swizzle: core matrix = 8 rows x 16 bytes (contiguous 128 bytes);
// sbo = bytes between 8-row groups, lbo = bytes between the two 8-wide K chunks of one K = 16 step; version 1 (sm_100)
__device__ __forceinline__ uint64_t sh_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo)
{
	return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) |
	       (1ull << 46);
}
// instruction descriptor (kind::f16): D = f32, A = B = bf16, both K-major, M = 128, N
__device__ __forceinline__ uint32_t sh_idesc(int n)
{
	return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(SH_M >> 4) << 24);
}
__device__ __forceinline__ void sh_mma(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
	asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
		     "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
		     "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
		     : "memory");
}
// arrive on `bar` when every tcgen05 operation issued so far by this thread has completed
__device__ __forceinline__ void sh_commit(uint64_t *bar)
{
	asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(tma_smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void sh_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void sh_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// 32 accumulator columns of this thread's row -> registers (asynchronous: sh_tmem_wait before the first use)
__device__ __forceinline__ void sh_tmem_ld32(uint32_t taddr, uint32_t (&r)[32])
{
	asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
		     "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
		     "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
		     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
		       "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
		       "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
		       "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
		     : "r"(taddr)
		     : "memory");
}
// wait for the outstanding tcgen05.ld; the registers are operands so that no use of them can be scheduled above the wait
__device__ __forceinline__ void sh_tmem_wait(uint32_t (&r)[32])
{
	asm volatile("tcgen05.wait::ld.sync.aligned;"
		     : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
		       "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]), "+r"(r[16]),
		       "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]), "+r"(r[24]),
		       "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
		     :
		     : "memory");
}
__device__ __forceinline__ void sh_bulk_load(void *smem_dst, const void *gsrc, uint32_t bytes, uint64_t *bar)
{
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(tma_smem_addr(smem_dst)),
		     "l"(gsrc), "r"(bytes), "r"(tma_smem_addr(bar))
		     : "memory");
}
// two floats -> packed bf16 pair (a at the lower address) and the residuals a - bf16(a), b - bf16(b)
__device__ __forceinline__ uint32_t sh_split2(float a, float b, float &ra, float &rb)
{
	uint32_t p;
	asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(p) : "f"(b), "f"(a));     // upper half <- first operand
	ra = a - __uint_as_float(p << 16);
	rb = b - __uint_as_float(p & 0xffff0000u);
	return p;
}
__device__ __forceinline__ uint32_t sh_pack2(float a, float b)
{
	uint32_t p;
	asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(p) : "f"(b), "f"(a));
	return p;
}
// 8 consecutive K values of one row -> one 16-byte piece of the hi operand and one of the lo operand
template <int NTERMS>
__device__ __forceinline__ void sh_store8(unsigned char *a_hi, unsigned char *a_lo, int kchunk, int m, const float (&v)[8])
{
	uint4 hi, lo;
	float r0, r1, r2, r3, r4, r5, r6, r7;
	hi.x = sh_split2(v[0], v[1], r0, r1);
	hi.y = sh_split2(v[2], v[3], r2, r3);
	hi.z = sh_split2(v[4], v[5], r4, r5);
	hi.w = sh_split2(v[6], v[7], r6, r7);
	*reinterpret_cast<uint4 *>(a_hi + kchunk * SH_ACHUNK + m * 16) = hi;
	if (NTERMS == 3) {
		lo.x = sh_pack2(r0, r1);
		lo.y = sh_pack2(r2, r3);
		lo.z = sh_pack2(r4, r5);
		lo.w = sh_pack2(r6, r7);
		*reinterpret_cast<uint4 *>(a_lo + kchunk * SH_ACHUNK + m * 16) = lo;
	}
}


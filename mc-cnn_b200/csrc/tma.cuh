// tma.cuh -- TMA (cp.async.bulk.tensor) + mbarrier helpers for the sm_100a kernels, and the host-side
// tensor-map encoder.  The driver entry point cuTensorMapEncodeTiled is resolved at run time through
// cudaGetDriverEntryPoint, so the library links against cudart only (no -lcuda).
//
// Private volumes of the fused pipeline are (D, H, ld) with a row pitch `ld` that is a multiple of 4
// floats: every global stride is then a multiple of 16 bytes, which is what a tensor map requires (the
// API-facing (D,H,W) tensors with W = 1226 are only 8-byte aligned per row and cannot be described).
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

// ---------------------------------------------------------------- host: tensor maps
// fp32 tensor of rank 2 or 3; dims/box innermost first; strides in BYTES for dims 1..rank-1.
// Out-of-bounds box elements are filled with zeros.  Returns 0 or a negative/positive error code.
int adc_tma_encode(CUtensorMap *map, const void *base, int rank, const uint64_t *dims, const uint64_t *strides_bytes,
		   const uint32_t *box);

// volume (D, H, ld) seen as a W x H x D tensor (columns >= W are out of bounds => zero fill)
static inline int adc_tma_encode_volume(CUtensorMap *map, const float *vol, int D, int H, int W, int ld, int box_w, int box_h)
{
	const uint64_t dims[3] = {(uint64_t)W, (uint64_t)H, (uint64_t)D};
	const uint64_t strides[2] = {(uint64_t)ld * 4, (uint64_t)ld * 4 * (uint64_t)H};
	const uint32_t box[3] = {(uint32_t)box_w, (uint32_t)box_h, 1};
	return adc_tma_encode(map, vol, 3, dims, strides, box);
}

// ---------------------------------------------------------------- device
#ifdef __CUDACC__
__device__ __forceinline__ uint32_t tma_smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(tma_smem_addr(bar)), "r"(count) : "memory");
}
// make mbarrier initialisation visible to the async proxy (TMA)
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// order generic-proxy shared-memory accesses before subsequent async-proxy (TMA) accesses
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes)
{
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(tma_smem_addr(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity)
{
	uint32_t done;
	asm volatile("{\n\t.reg .pred p;\n\t"
		     "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
		     "selp.b32 %0, 1, 0, p;\n\t}"
		     : "=r"(done)
		     : "r"(tma_smem_addr(bar)), "r"(parity)
		     : "memory");
	return done != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
	while (!mbar_try_wait(bar, parity)) {
	}
}

// box of a rank-3 tensor -> shared memory, completion counted in bytes on `bar`.  c0 (innermost coordinate) must be a
// multiple of 16 bytes / element size (measured: any other start raises 'illegal instruction'), smem_dst 128-byte aligned
__device__ __forceinline__ void tma_load_3d(void *smem_dst, const CUtensorMap *map, int c0, int c1, int c2, uint64_t *bar)
{
	asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
		     :
		     : "r"(tma_smem_addr(smem_dst)), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(tma_smem_addr(bar))
		     : "memory");
}
__device__ __forceinline__ void tma_load_2d(void *smem_dst, const CUtensorMap *map, int c0, int c1, uint64_t *bar)
{
	asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
		     :
		     : "r"(tma_smem_addr(smem_dst)), "l"(map), "r"(c0), "r"(c1), "r"(tma_smem_addr(bar))
		     : "memory");
}
// shared memory -> box of a rank-3 tensor (bulk async-group completion); same alignment rules as the loads
__device__ __forceinline__ void tma_store_3d(const CUtensorMap *map, const void *smem_src, int c0, int c1, int c2)
{
	asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%1, %2, %3}], [%4];"
		     :
		     : "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(tma_smem_addr(smem_src))
		     : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// all bulk stores of this thread have finished READING shared memory (the buffer may be reused / the CTA may exit)
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }

__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap *map)
{
	asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
#endif

// scorer_head.cu -- the accurate ('slow') architecture's scorer head as ONE fused tensor-core kernel.
//
// Reference (main.lua:958-984, 688-695; SpatialConvolution1_fw.lua:11-31): for every disparity d the tower outputs are
// sliced (l = left[:, :, d:], r = right[:, :, :W-d]), stacked along channels (2 fm) and pushed through net_te2 =
// l2 x [per-pixel W x + b (cuBLAS addmm), ReLU], then W5 x + b5 (nh2 -> 1) and a sigmoid: 228 x 5 GEMM launches plus 228
// slice / concat copies per direction, activations through HBM between them.  The (H, W-d) result is the matching cost of
// disparity d, stored at vol[d, :, d:] (direction -1) and vol[d, :, :W-d] (direction +1): the same numbers, so they are
// computed ONCE here and written to both volumes.
//
// B200 design.  A tile is 128 consecutive pixels of one image row at one disparity = 128 rows of the per-pixel MLP; the
// whole chain runs on it without leaving the SM:
//   * tcgen05.mma (kind::f16, bf16 operands, fp32 accumulation), M = 128, N = nh2 (384 = one 256- and one 128-wide
//     instruction per K = 16 step), accumulator in TMEM (384 of 512 columns), issued by one thread;
//   * fp32 parity through a bf16 SPLIT: x = hi + lo (two bf16), x . w ~= hi.whi + lo.whi + hi.wlo -- three MMAs per K step
//     into the same accumulator (the dropped lo.wlo term is 2^-16 relative): ~1e-6 relative per layer, the north star's
//     1e-4 holds with two orders of margin.  nterms = 1 (plain bf16, ~1e-3) is selectable for speed;
//   * activations stay on chip: the A operand of a layer (128 x K, hi and lo, 196 KB at K = 384) lives in shared memory in
//     the canonical K-major no-swizzle core-matrix layout [K/8][128 rows][16 bytes]; the epilogue of a layer (4 warps,
//     thread = row = TMEM lane: tcgen05.ld 32 columns at a time, + bias, ReLU, split, 16-byte stores) writes the next
//     layer's A operand in place;
//   * weights stream: pre-split once into per-K-step slabs [2 chunks][N][8 bf16] (exactly the shared-memory image the MMA
//     wants), fetched by ONE thread with cp.async.bulk into a small ring (full / empty mbarriers, tcgen05.commit frees a
//     slot); they are L2-resident (2.1 MB for kitti) and shared by all CTAs;
//   * layer 1 takes its rows straight from the (fm, H, W) tower outputs (left pixel x, right pixel x - d: coalesced loads
//     along x), the last hidden layer's epilogue folds the nh2 -> 1 product, bias and sigmoid and stores the two volumes.
// A CTA owns (image row, 128-pixel tile) and loops over the disparities.  Roles: warps 0-7 build / read the tile (two
// threads per row, one per column half; warp w may touch the TMEM lanes of quadrant w % 4), warp 8 streams weights,
// warp 9 allocates TMEM and issues the MMAs.
//
// Roofline: tensor-bound.  Per valid (pixel, d): 2 * (2 fm * nh2 + (l2 - 1) * nh2^2) flop = 1.06 MFLOP for kitti
// (fm 112, nh2 384, l2 4), x 3 for the split.
#include <stdlib.h>
#include <string.h>

#include "umma.cuh"

namespace {

constexpr int SH_NMAX = 384, SH_KMAX = 384, SH_LMAX = 4;
constexpr int SH_SLAB = 2 * SH_NMAX * 16;    // ring slot: one K = 16 step of B, [2 chunks][N][8 bf16]
constexpr int SH_ABYTES = (SH_KMAX / 8) * SH_ACHUNK;
constexpr int SH_NFEED = 256;                // tile builders / epilogue: 2 threads per row (column halves)
constexpr int SH_NT = SH_NFEED + 64;         // + weight producer warp + MMA warp

template <int NTERMS>
struct SHCfg {
	static constexpr int NSLOT = NTERMS == 3 ? 2 : 10;
	static constexpr int OFF_AHI = 0;
	static constexpr int OFF_ALO = SH_ABYTES;
	static constexpr int OFF_RING = NTERMS == 3 ? 2 * SH_ABYTES : SH_ABYTES;
	static constexpr int OFF_CONST = OFF_RING + NSLOT * SH_SLAB;            // floats: bias[LMAX][NMAX], w5[NMAX], b5
	static constexpr int NCONST = SH_LMAX * SH_NMAX + SH_NMAX + 4;
	static constexpr int OFF_BAR = OFF_CONST + NCONST * 4;                   // full[NSLOT], empty[NSLOT], a_ready, acc_full
	static constexpr int OFF_TPTR = OFF_BAR + (2 * NSLOT + 2) * 8;
	static constexpr int OFF_PART = OFF_TPTR + 16;                           // [128] partial dot products of column half 1
	static constexpr int SMEM = OFF_PART + SH_M * 4;
	static_assert(SMEM <= 232448, "shared memory budget");
};

struct SHParams {
	const float *featL, *featR;
	float *volL, *volR;
	const unsigned char *wslabs;     // per layer, per K step: [hi slab][lo slab], slab = [2][N][8] bf16
	const float *consts;             // bias of the hidden layers [L][NMAX], w5[NMAX], b5
	long slab_off[SH_LMAX];          // byte offset of layer l's first slab
	int nk[SH_LMAX];                 // K steps (of 16) of layer l
	int fm, N, L, H, W, D;
};

template <int NTERMS>
__global__ void __launch_bounds__(SH_NT, 1)
scorer_head_kernel(const SHParams p)
{
	using C = SHCfg<NTERMS>;
	constexpr int NSLOT = C::NSLOT;
	extern __shared__ __align__(128) unsigned char sh_smem[];
	unsigned char *a_hi = sh_smem + C::OFF_AHI, *a_lo = sh_smem + C::OFF_ALO;
	unsigned char *ring = sh_smem + C::OFF_RING;
	float *cst = reinterpret_cast<float *>(sh_smem + C::OFF_CONST);
	uint64_t *bar_full = reinterpret_cast<uint64_t *>(sh_smem + C::OFF_BAR);
	uint64_t *bar_empty = bar_full + NSLOT;
	uint64_t *bar_a = bar_empty + NSLOT;       // the A operand of the next layer is in shared memory
	uint64_t *bar_acc = bar_a + 1;             // the accumulator of the current layer is complete
	uint32_t *tptr = reinterpret_cast<uint32_t *>(sh_smem + C::OFF_TPTR);
	float *part = reinterpret_cast<float *>(sh_smem + C::OFF_PART);

	const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
	const int x0 = blockIdx.x * SH_M, y = blockIdx.y;
	const int N = p.N, L = p.L, W = p.W;
	const long HW = (long)p.H * W;
	// disparities with at least one valid pixel in this tile: d <= x (adcensus / main.lua:967-968), d < D
	const int xl = min(x0 + SH_M, W) - 1;
	const int nd = min(p.D, xl + 1);
	const uint32_t slab_bytes = (uint32_t)N * 32u;

	if (tid == 0) {
		for (int s = 0; s < NSLOT; s++) {
			mbar_init(&bar_full[s], 1);
			mbar_init(&bar_empty[s], 1);
		}
		mbar_init(bar_a, SH_NFEED);
		mbar_init(bar_acc, 1);
		mbar_fence_init();
	}
	for (int i = tid; i < C::NCONST; i += SH_NT) cst[i] = p.consts[i];
	if (warp == 9) {                               // TMEM: 512 columns (the accumulator needs N <= 384)
		asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tma_smem_addr(tptr)), "r"(512) : "memory");
		asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
	}
	sh_fence_before();
	__syncthreads();
	sh_fence_after();
	const uint32_t tbase = *tptr;

	if (warp == 8) {
		// ------------------------------------------------------------ weight producer
		if (lane == 0) {
			unsigned it = 0;
			for (int t = 0; t < nd; t++)
				for (int l = 0; l < L; l++) {
					const unsigned char *src = p.wslabs + p.slab_off[l];
					for (int ks = 0; ks < p.nk[l]; ks++)
						for (int part = 0; part < (NTERMS == 3 ? 2 : 1); part++, it++) {
							const int s = it % NSLOT;
							if (it >= (unsigned)NSLOT) sh_wait(&bar_empty[s], ((it / NSLOT) - 1) & 1);
							mbar_arrive_expect_tx(&bar_full[s], slab_bytes);
							sh_bulk_load(ring + s * SH_SLAB, src + ((long)ks * 2 + part) * slab_bytes, slab_bytes, &bar_full[s]);
						}
				}
		}
	} else if (warp == 9) {
		// ------------------------------------------------------------ MMA issuer
		if (lane == 0) {
			const uint32_t a_hi_s = tma_smem_addr(a_hi), a_lo_s = tma_smem_addr(a_lo), ring_s = tma_smem_addr(ring);
			const uint32_t lbo_b = (uint32_t)N * 16u;
			const int n0 = N >= 256 ? 256 : N, n1 = N - n0;      // one or two instructions per K step
			const uint32_t id0 = sh_idesc(n0), id1 = n1 ? sh_idesc(n1) : 0u;
			unsigned it = 0, use = 0;
			for (int t = 0; t < nd; t++)
				for (int l = 0; l < L; l++, use++) {
					sh_wait(bar_a, use & 1);
					sh_fence_after();
					for (int ks = 0; ks < p.nk[l]; ks++) {
						const uint64_t da_hi = sh_desc(a_hi_s + ks * 2 * SH_ACHUNK, SH_ACHUNK, 128);
						const uint64_t da_lo = sh_desc(a_lo_s + ks * 2 * SH_ACHUNK, SH_ACHUNK, 128);
						{   // w_hi slab: a_hi . w_hi (+ a_lo . w_hi)
							const int s = it % NSLOT;
							sh_wait(&bar_full[s], (it / NSLOT) & 1);
							sh_fence_after();
							const uint32_t b_s = ring_s + s * SH_SLAB;
							const uint64_t db0 = sh_desc(b_s, lbo_b, 128), db1 = sh_desc(b_s + n0 * 16, lbo_b, 128);
							sh_mma(tbase, da_hi, db0, id0, ks > 0);
							if (n1) sh_mma(tbase + n0, da_hi, db1, id1, ks > 0);
							if (NTERMS == 3) {
								sh_mma(tbase, da_lo, db0, id0, 1);
								if (n1) sh_mma(tbase + n0, da_lo, db1, id1, 1);
							}
							sh_commit(&bar_empty[s]);
							it++;
						}
						if (NTERMS == 3) {   // w_lo slab: a_hi . w_lo
							const int s = it % NSLOT;
							sh_wait(&bar_full[s], (it / NSLOT) & 1);
							sh_fence_after();
							const uint32_t b_s = ring_s + s * SH_SLAB;
							const uint64_t db0 = sh_desc(b_s, lbo_b, 128), db1 = sh_desc(b_s + n0 * 16, lbo_b, 128);
							sh_mma(tbase, da_hi, db0, id0, 1);
							if (n1) sh_mma(tbase + n0, da_hi, db1, id1, 1);
							sh_commit(&bar_empty[s]);
							it++;
						}
					}
					sh_commit(bar_acc);
				}
		}
	} else {
		// ------------------------------------------------------------ tile builders / epilogue
		// thread = (row m = TMEM lane, column half hf): warp w covers lanes 32 (w % 4) .. + 31, half w / 4
		const int m = (warp & 3) * 32 + lane, hf = warp >> 2;
		const int x = x0 + m;
		const int fm = p.fm, nc1 = 2 * fm / 8, ncl = fm / 8;
		const int kc_a = hf ? (nc1 + 1) / 2 : 0, kc_b = hf ? nc1 : (nc1 + 1) / 2;     // layer-1 K chunks of this half
		const int cn = N / 2, ca = hf * cn;                                            // accumulator columns of this half
		const uint32_t trow = tbase + ((uint32_t)((warp & 3) * 32) << 16);
		const float *w5 = cst + SH_LMAX * SH_NMAX;
		const float b5 = w5[SH_NMAX];
		unsigned use = 0;
		for (int d = 0; d < nd; d++) {
			const bool rowok = x < W && x >= d;
			// layer 1 input: [left pixel x ; right pixel x - d], main.lua:967-970, four K chunks (32 loads) in flight
			{
				const float *pl = p.featL + (long)y * W + (rowok ? x : 0);
				const float *pr = p.featR + (long)y * W + (rowok ? x - d : 0);
				for (int kc0 = kc_a; kc0 < kc_b; kc0 += 4) {
					float v[4][8];
#pragma unroll
					for (int u = 0; u < 4; u++) {
						const int kc = kc0 + u;
						const float *src = kc < ncl ? pl + (long)(8 * kc) * HW : pr + (long)(8 * (kc - ncl)) * HW;
#pragma unroll
						for (int e = 0; e < 8; e++) v[u][e] = (rowok && kc < kc_b) ? __ldg(src + (long)e * HW) : 0.0f;
					}
#pragma unroll
					for (int u = 0; u < 4; u++)
						if (kc0 + u < kc_b) sh_store8<NTERMS>(a_hi, a_lo, kc0 + u, m, v[u]);
				}
			}
			fence_proxy_async_smem();              // generic stores -> visible to the tensor core's (async proxy) reads
			sh_fence_before();
			sh_arrive(bar_a);
			for (int l = 0; l < L; l++, use++) {
				sh_wait(bar_acc, use & 1);
				sh_fence_after();
				const float *bias = cst + l * SH_NMAX;
				// 32 columns at a time, the next 32 already in flight while these are processed
				uint32_t ra[32], rb[32];
				sh_tmem_ld32(trow + ca, ra);
				if (l + 1 < L) {
					// hidden layer: + bias, ReLU (SpatialConvolution1_fw.lua:21-27, cudnn.ReLU), next layer's operand
					for (int c0 = ca; c0 < ca + cn; c0 += 64) {
						sh_tmem_wait(ra);
						if (c0 + 32 < ca + cn) sh_tmem_ld32(trow + c0 + 32, rb);
#pragma unroll
						for (int g = 0; g < 4; g++) {
							float v[8];
#pragma unroll
							for (int e = 0; e < 8; e++) v[e] = fmaxf(__uint_as_float(ra[8 * g + e]) + bias[c0 + 8 * g + e], 0.0f);
							sh_store8<NTERMS>(a_hi, a_lo, c0 / 8 + g, m, v);
						}
						if (c0 + 32 < ca + cn) {
							sh_tmem_wait(rb);
							if (c0 + 64 < ca + cn) sh_tmem_ld32(trow + c0 + 64, ra);
#pragma unroll
							for (int g = 0; g < 4; g++) {
								float v[8];
#pragma unroll
								for (int e = 0; e < 8; e++)
									v[e] = fmaxf(__uint_as_float(rb[8 * g + e]) + bias[c0 + 32 + 8 * g + e], 0.0f);
								sh_store8<NTERMS>(a_hi, a_lo, (c0 + 32) / 8 + g, m, v);
							}
						}
					}
					fence_proxy_async_smem();
					sh_fence_before();
					sh_arrive(bar_a);
				} else {
					// last hidden layer folded with the nh2 -> 1 layer and the sigmoid (main.lua:693-694)
					float acc = 0.0f;
					for (int c0 = ca; c0 < ca + cn; c0 += 64) {
						sh_tmem_wait(ra);
						if (c0 + 32 < ca + cn) sh_tmem_ld32(trow + c0 + 32, rb);
#pragma unroll
						for (int e = 0; e < 32; e++)
							acc = fmaf(fmaxf(__uint_as_float(ra[e]) + bias[c0 + e], 0.0f), w5[c0 + e], acc);
						if (c0 + 32 < ca + cn) {
							sh_tmem_wait(rb);
							if (c0 + 64 < ca + cn) sh_tmem_ld32(trow + c0 + 64, ra);
#pragma unroll
							for (int e = 0; e < 32; e++)
								acc = fmaf(fmaxf(__uint_as_float(rb[e]) + bias[c0 + 32 + e], 0.0f), w5[c0 + 32 + e], acc);
						}
					}
					sh_fence_before();             // the accumulator has been read: the next tile's MMAs may overwrite it
					if (hf) part[m] = acc;
					asm volatile("bar.sync 1, %0;" ::"n"(SH_NFEED) : "memory");
					if (!hf) {
						const float z = (acc + part[m]) + b5;
						const float s = 1.0f / (1.0f + expf(-z));
						if (rowok) {
							const long o = ((long)d * p.H + y) * W;
							if (p.volL) p.volL[o + x] = s;           // main.lua:976, direction -1: columns d ..
							if (p.volR) p.volR[o + x - d] = s;       //                direction +1: columns .. W - d
						}
					}
					asm volatile("bar.sync 1, %0;" ::"n"(SH_NFEED) : "memory");   // part[] free for the next tile
				}
			}
		}
	}
	sh_fence_before();
	__syncthreads();
	if (warp == 9) {
		sh_fence_after();
		asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(512) : "memory");
	}
}

// weights (N x K fp32, row-major) -> slabs: per K step [hi: [2][N][8] bf16][lo: [2][N][8] bf16]
__global__ void sh_prep_kernel(const float *__restrict__ w, unsigned short *__restrict__ slabs, int N, int K)
{
	const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= (long)N * K) return;
	const int n = (int)(i / K), k = (int)(i % K);
	const float v = w[i];
	const unsigned hb = __float_as_uint(v);
	// round to nearest even bf16 (finite weights)
	const unsigned hi = (hb + 0x7fffu + ((hb >> 16) & 1u)) >> 16;
	const float r = v - __uint_as_float(hi << 16);
	const unsigned rb = __float_as_uint(r);
	const unsigned lo = (rb + 0x7fffu + ((rb >> 16) & 1u)) >> 16;
	const int ks = k / 16, j = (k / 8) & 1, e = k & 7;
	const long slab = (long)N * 16;                       // bf16 elements per slab
	const long base = (long)ks * 2 * slab + ((long)j * N + n) * 8 + e;
	slabs[base] = (unsigned short)hi;
	slabs[base + slab] = (unsigned short)lo;
}

}  // namespace

struct mccnn_scorer_head {
	int fm, nh2, l2, device;
	unsigned char *wslabs;
	float *consts;
	long slab_off[SH_LMAX];
	int nk[SH_LMAX];
};

// W[i] (out_i x in_i, row-major) / b[i] (out_i) for i = 0 .. l2 (the last one is the nh2 -> 1 layer): DEVICE pointers,
// the layout of net_te2's SpatialConvolution1_fw modules (main.lua:688-695).
extern "C" int mccnn_scorer_head_create(mccnn_scorer_head **out, int fm, int nh2, int l2, const float *const *W, const float *const *b,
					 int device, adcensus_stream_t stream)
{
	if (!out || !W || !b || fm < 8 || nh2 < 128 || l2 < 1) return ADCENSUS_EINVAL;
	if (l2 > SH_LMAX || nh2 > SH_NMAX || (nh2 % 128) || (fm % 8) || 2 * fm > SH_KMAX) return ADCENSUS_ELIMIT;
	int prev = 0;
	cudaGetDevice(&prev);
	cudaSetDevice(device);
	cudaStream_t s = adc_stream(stream);
	mccnn_scorer_head *h = (mccnn_scorer_head *)calloc(1, sizeof(*h));
	if (!h) return ADCENSUS_EINVAL;
	h->fm = fm; h->nh2 = nh2; h->l2 = l2; h->device = device;
	int rc = 0;
	long total = 0;
	for (int l = 0; l < l2; l++) {
		const int K = l == 0 ? 2 * fm : nh2;
		h->nk[l] = (K + 15) / 16;
		h->slab_off[l] = total;
		total += (long)h->nk[l] * 2 * nh2 * 32;
	}
	rc = (int)cudaMalloc((void **)&h->wslabs, total);
	if (!rc) rc = (int)cudaMemsetAsync(h->wslabs, 0, total, s);            // K padding of layer 1 (2 fm not a multiple of 16)
	if (!rc) rc = (int)cudaMalloc((void **)&h->consts, SHCfg<3>::NCONST * sizeof(float));
	if (!rc) rc = (int)cudaMemsetAsync(h->consts, 0, SHCfg<3>::NCONST * sizeof(float), s);
	for (int l = 0; l < l2 && !rc; l++) {
		const int K = l == 0 ? 2 * fm : nh2;
		const long n = (long)nh2 * K;
		sh_prep_kernel<<<adc_div_up(n, 256), 256, 0, s>>>(W[l], (unsigned short *)(h->wslabs + h->slab_off[l]), nh2, K);
		rc = (int)cudaPeekAtLastError();
		if (!rc) rc = (int)cudaMemcpyAsync(h->consts + l * SH_NMAX, b[l], nh2 * sizeof(float), cudaMemcpyDeviceToDevice, s);
	}
	if (!rc) rc = (int)cudaMemcpyAsync(h->consts + SH_LMAX * SH_NMAX, W[l2], nh2 * sizeof(float), cudaMemcpyDeviceToDevice, s);
	if (!rc) rc = (int)cudaMemcpyAsync(h->consts + SH_LMAX * SH_NMAX + SH_NMAX, b[l2], sizeof(float), cudaMemcpyDeviceToDevice, s);
	if (!rc) rc = (int)cudaStreamSynchronize(s);
	cudaSetDevice(prev);
	if (rc) {
		cudaFree(h->wslabs);
		cudaFree(h->consts);
		free(h);
		return rc;
	}
	*out = h;
	return 0;
}

extern "C" void mccnn_scorer_head_destroy(mccnn_scorer_head *h)
{
	if (!h) return;
	cudaFree(h->wslabs);
	cudaFree(h->consts);
	free(h);
}

// featL / featR: (fm, H, W) tower outputs of the left / right image; volL / volR: (D, H, W), either may be NULL.  Entries
// with x - d < 0 (volL) / x + d >= W (volR) are NOT written (the caller pre-fills NaN like main.lua:962, and applies
// fix_border afterwards, :981).  nterms: 3 = bf16 split (fp32-grade, default), 1 = plain bf16.
extern "C" int mccnn_scorer_head_forward(const mccnn_scorer_head *h, const float *featL, const float *featR, float *volL, float *volR,
					  int H, int W, int D, int nterms, adcensus_stream_t stream)
{
	if (!h || !featL || !featR || (!volL && !volR) || H < 1 || W < 1 || D < 1 || (nterms != 1 && nterms != 3)) return ADCENSUS_EINVAL;
	if (H > 65535) return ADCENSUS_ELIMIT;
	SHParams p;
	memset(&p, 0, sizeof(p));
	p.featL = featL; p.featR = featR; p.volL = volL; p.volR = volR;
	p.wslabs = h->wslabs; p.consts = h->consts;
	for (int l = 0; l < h->l2; l++) {
		p.slab_off[l] = h->slab_off[l];
		p.nk[l] = h->nk[l];
	}
	p.fm = h->fm; p.N = h->nh2; p.L = h->l2; p.H = H; p.W = W; p.D = D;
	cudaStream_t s = adc_stream(stream);
	const dim3 grid(adc_div_up(W, SH_M), H);
	static bool attr3[64] = {false}, attr1[64] = {false};
	int dev = 0;
	cudaGetDevice(&dev);
	if (nterms == 3) {
		if (!attr3[dev & 63]) {
			ADC_CUDA(cudaFuncSetAttribute(scorer_head_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, SHCfg<3>::SMEM));
			attr3[dev & 63] = true;
		}
		scorer_head_kernel<3><<<grid, SH_NT, SHCfg<3>::SMEM, s>>>(p);
	} else {
		if (!attr1[dev & 63]) {
			ADC_CUDA(cudaFuncSetAttribute(scorer_head_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, SHCfg<1>::SMEM));
			attr1[dev & 63] = true;
		}
		scorer_head_kernel<1><<<grid, SH_NT, SHCfg<1>::SMEM, s>>>(p);
	}
	ADC_CHECK_LAUNCH();
	return 0;
}

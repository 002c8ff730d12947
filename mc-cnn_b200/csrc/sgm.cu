// sgm.cu -- adcensus.sgm2 for sm_100a.
//
// Replaces adcensus.cu:535-697: the reference launches one kernel per scan step
// (2*(W+H) = 3192 launches at 370x1226), each block re-reading the previous
// step's line state from global `tmp` and tree-reducing the min over D in shared
// memory.  Here one launch does a whole direction:
//   * ONE WARP PER SCANLINE keeps the line state L_r(p-r, .) in registers for the
//     entire scan; lane l owns the K consecutive disparities l*K .. l*K+K-1;
//   * the min over D is a K-element register tree + one `redux.sync.min.f32`
//     (CREDUX, sm_100a) instead of a shared-memory tree with 9 barriers;
//   * the d-1 / d+1 neighbours are registers (one shuffle each at lane borders);
//   * the cost vectors of the next PF pixels stream into a per-warp shared-memory
//     ring with cp.async (16-byte granules, each lane fetching exactly the
//     elements it will consume, so no barrier is needed), which keeps ~PF*2 KB per
//     warp in flight and hides HBM latency behind the serial recurrence;
//   * the P1/P2 selection (adcensus.cu:586-605) is table driven: a pre-pass
//     classifies |I(p) - I(p-r)| against tau_so once per image / scan axis into
//     byte tables (padded by D columns of the out-of-image class, D2 = 10), so a
//     step needs one byte per disparity instead of two image loads and two
//     bounds checks.
//
// Arithmetic is the reference's, expression for expression (adds, fminf, IEEE
// divisions; nothing contractible), and `out` is accumulated in the reference's
// direction order (right, left, down, up) => bit-identical results for volumes
// whose valid disparities form a prefix per pixel (what StereoJoin / ad / census
// produce).  Padding slots d >= D are carried as NaN, which fminf ignores exactly
// like the reference's `d + 1 < size3` guard.
//
// Layout: input/output (H,W,D) like the reference (INDEX, adcensus.cu:531-533).
// Algorithmic traffic per call: 4 passes x (read input + read-modify-write
// output) = 12V bytes (V = 4*D*H*W); 11V when the caller guarantees a zeroed
// output (pipeline), because the first pass then skips the read.
#include "common.cuh"

namespace {

struct SgmParams {
	float pi1, pi2, tau_so, alpha1, q1, q2;
	int direction;
	// band support: the volume (H,W,D) handed to a pass may be a row band or a column band of the
	// image; its pixel (y, x) is image pixel (yoff + y, xoff + x) of the Ht x Wt image the class
	// tables were built from.  Bands never cut a scanline of the pass they are used for.
	int Ht, Wt, yoff, xoff;
};

// ---------------------------------------------------------------- penalty class tables
// class of a colour difference against tau_so: 0 (< tau), 2 (> tau), 1 otherwise (== tau or NaN):
// exactly the three branches of adcensus.cu:596-605.
__device__ __forceinline__ uint8_t sgm_class(float diff, float tau)
{
	return diff < tau ? 0 : (diff > tau ? 2 : 1);
}

// tab layout: 4 planes [h0, v0, h1, v1], each H rows of pitch Wp = W + 2*pad bytes, image column
// j at byte pad + j.  h: |I[y][j] - I[y][j-1]|, v: |I[y][j] - I[y-1][j]|; entries whose
// neighbour is outside the image, and the padding, hold the class of D2 = 10 (adcensus.cu:591).
// Plane 0/1 from x0 (D1, always in range where used), plane 2/3 from x1 (D2).
__global__ void sgm_class_kernel(const float *__restrict__ x0, const float *__restrict__ x1, uint8_t *__restrict__ tab,
				 int H, int W, int pad, float tau)
{
	const int Wp = W + 2 * pad;
	const long plane = (long)H * Wp;
	long id = (long)blockIdx.x * blockDim.x + threadIdx.x;
	if (id >= plane) return;
	int y = (int)(id / Wp), j = (int)(id % Wp) - pad;
	const uint8_t oob = sgm_class(10.0f, tau);
	uint8_t h0 = oob, v0 = oob, h1 = oob, v1 = oob;
	if (j >= 0 && j < W) {
		long p = (long)y * W + j;
		if (j >= 1) {
			h0 = sgm_class(fabsf(x0[p] - x0[p - 1]), tau);
			h1 = sgm_class(fabsf(x1[p] - x1[p - 1]), tau);
		}
		if (y >= 1) {
			v0 = sgm_class(fabsf(x0[p] - x0[p - W]), tau);
			v1 = sgm_class(fabsf(x1[p] - x1[p - W]), tau);
		}
	}
	tab[id] = h0;
	tab[plane + id] = v0;
	tab[2 * plane + id] = h1;
	tab[3 * plane + id] = v1;
}

// ---------------------------------------------------------------- helpers
__device__ __forceinline__ float warp_min_f32(float v)
{
	float r;
	asm("redux.sync.min.f32 %0, %1, 0xffffffff;" : "=f"(r) : "f"(v));  // NaN inputs are skipped
	return r;
}

__device__ __forceinline__ void cp_async16(float *smem_dst, const float *gsrc)
{
	asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc));
}
__device__ __forceinline__ void cp_async4(float *smem_dst, const float *gsrc)
{
	asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc));
}

template <int K, bool VEC>
__device__ __forceinline__ void issue_vec(float *slot, const float *g, int dbase, int D)
{
	if constexpr (VEC) {
#pragma unroll
		for (int k = 0; k < K; k += 4)
			if (dbase + k < D) cp_async16(slot + k, g + k);
	} else {
#pragma unroll
		for (int k = 0; k < K; k++)
			if (dbase + k < D) cp_async4(slot + k, g + k);
	}
}

template <int K>
__device__ __forceinline__ void read_slot(float (&r)[K], const float *slot)
{
	if constexpr (K % 4 == 0) {
#pragma unroll
		for (int k = 0; k < K; k += 4) {
			float4 v = *reinterpret_cast<const float4 *>(slot + k);
			r[k] = v.x; r[k + 1] = v.y; r[k + 2] = v.z; r[k + 3] = v.w;
		}
	} else {
#pragma unroll
		for (int k = 0; k < K; k++) r[k] = slot[k];
	}
}

template <int K, bool VEC>
__device__ __forceinline__ void store_vec(const float (&r)[K], float *p, int dbase, int D)
{
	if constexpr (VEC) {
#pragma unroll
		for (int k = 0; k < K; k += 4)
			if (dbase + k < D) *reinterpret_cast<float4 *>(p + k) = make_float4(r[k], r[k + 1], r[k + 2], r[k + 3]);
	} else {
#pragma unroll
		for (int k = 0; k < K; k++)
			if (dbase + k < D) p[k] = r[k];
	}
}

// ---------------------------------------------------------------- one scan direction
// SD: 0 right, 1 left, 2 down, 3 up (adcensus.cu:541-565).  ZERO: output known to be 0 on entry.
// NR: scanlines per warp.  A horizontal scan has only H (370) lines of W (1226) strictly serial
// steps, one warp per SM sub-partition; interleaving NR = 2 independent lines in one warp fills the
// dependent-issue bubbles of the recurrence (min tree -> CREDUX -> fminf chain).
template <int K, bool VEC, int SD, bool ZERO, int PF, int WPB, int NR>
__global__ void __launch_bounds__(32 * WPB)
sgm_pass_kernel(const uint8_t *__restrict__ tab, const float *__restrict__ in, float *__restrict__ out,
		int H, int W, int D, int pad, SgmParams prm)
{
	extern __shared__ __align__(16) float sgm_smem[];
	constexpr int VSZ = 32 * K;                    // floats per cost vector slot (padded to 32*K)
	constexpr int NV = ZERO ? 1 : 2;               // ring holds `in` (and `out` unless ZERO)
	const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
	const int line0 = (blockIdx.x * WPB + wib) * NR;
	const int nlines = SD < 2 ? H : W;
	const int nsteps = SD < 2 ? W : H;
	if (line0 >= nlines) return;                   // whole warp

	constexpr int dx = SD == 0 ? 1 : (SD == 1 ? -1 : 0);
	constexpr int dy = SD == 2 ? 1 : (SD == 3 ? -1 : 0);
	const int dbase = lane * K;
	float *ring = sgm_smem + (size_t)wib * NR * PF * NV * VSZ + dbase;   // this lane's K floats of slot 0

	// adcensus.cu:595-605 and :609/:612, same expressions
	const float P1f = prm.pi1, P2f = prm.pi2;
	const float P1s = prm.pi1 / (prm.q1 * prm.q2), P2s = prm.pi2 / (prm.q1 * prm.q2);
	const float P1m = prm.pi1 / prm.q1, P2m = prm.pi2 / prm.q1;
	const float P1f_a = P1f / prm.alpha1, P1s_a = P1s / prm.alpha1, P1m_a = P1m / prm.alpha1;

	// class tables: D1 from plane (SD<2 ? 0 : 1), D2 from plane (SD<2 ? 2 : 3).  The stored
	// difference at (y, j) pairs pixel j with its left / upper neighbour, so scans that look
	// right / down (dx = -1, dy = -1) read the entry one further.
	const int Wp = prm.Wt + 2 * pad;
	const long plane = (long)prm.Ht * Wp;
	const uint8_t *t1 = tab + (SD < 2 ? 0 : 1) * plane + pad + prm.xoff;
	const uint8_t *t2 = tab + (SD < 2 ? 2 : 3) * plane + pad + prm.xoff;
	constexpr int tshift_x = dx < 0 ? 1 : 0;
	constexpr int tshift_y = dy < 0 ? 1 : 0;
	const int ddir = prm.direction;
	const long pix_step = (long)(dy * W + dx) * D;

	bool live[NR];
	int x[NR], y[NR];
	long base[NR];
	float *rg[NR];
#pragma unroll
	for (int r = 0; r < NR; r++) {
		const int line = line0 + r;
		live[r] = line < nlines;                   // warp-uniform
		const int ln = live[r] ? line : line0;
		x[r] = SD == 0 ? 0 : (SD == 1 ? W - 1 : ln);
		y[r] = SD == 2 ? 0 : (SD == 3 ? H - 1 : ln);
		base[r] = ((long)y[r] * W + x[r]) * D + dbase;
		rg[r] = ring + r * PF * NV * VSZ;
	}

	// NaN in the padding slots (never overwritten), then fill the ring
#pragma unroll
	for (int s = 0; s < NR * PF * NV; s++)
#pragma unroll
		for (int k = 0; k < K; k++) ring[s * VSZ + k] = adc_nan();
#pragma unroll
	for (int u = 0; u < PF; u++) {
#pragma unroll
		for (int r = 0; r < NR; r++)
			if (live[r] && u < nsteps) {
				issue_vec<K, VEC>(rg[r] + (u * NV) * VSZ, in + base[r] + u * pix_step, dbase, D);
				if (!ZERO) issue_vec<K, VEC>(rg[r] + (u * NV + 1) * VSZ, out + base[r] + u * pix_step, dbase, D);
			}
		asm volatile("cp.async.commit_group;");
	}

	// penalty classes are fetched one step ahead (the row of the table changes every step of a
	// vertical scan, so these loads miss L1; fetched just in time they were the top stall)
	uint8_t c1n[NR], c2n[NR][K];
	auto fetch_classes = [&](int r, int xs, int ys) {
		const int ty = min(max(ys + prm.yoff + tshift_y, 0), prm.Ht - 1);   // image row of the stored difference
		const uint8_t *q = t2 + (long)ty * Wp + xs + tshift_x + dbase * ddir;  // D2 classes (:588-594)
		c1n[r] = __ldg(t1 + (long)ty * Wp + xs + tshift_x);                     // D1 class (:587)
#pragma unroll
		for (int k = 0; k < K; k++) c2n[r][k] = __ldg(q + k * ddir);
	};
#pragma unroll
	for (int r = 0; r < NR; r++) fetch_classes(r, x[r] + dx, y[r] + dy);      // for step 1

	float prev[NR][K];
	for (int s0 = 0; s0 < nsteps; s0 += PF) {
#pragma unroll
		for (int u = 0; u < PF; u++) {
			const int s = s0 + u;
			if (s >= nsteps) break;
			asm volatile("cp.async.wait_group %0;" ::"n"(PF - 1));
#pragma unroll
			for (int r = 0; r < NR; r++) {
				if (!live[r]) continue;
				float cin[K], cout[K];
				read_slot<K>(cin, rg[r] + (u * NV) * VSZ);
				if (!ZERO) read_slot<K>(cout, rg[r] + (u * NV + 1) * VSZ);

				float val[K];
				if (s == 0) {                                   // adcensus.cu:567-572
#pragma unroll
					for (int k = 0; k < K; k++) val[k] = cin[k];
				} else {
					float mt[K];
#pragma unroll
					for (int k = 0; k < K; k++) mt[k] = prev[r][k];
#pragma unroll
					for (int w = K / 2; w > 0; w >>= 1)
#pragma unroll
						for (int k = 0; k < w; k++) mt[k] = fminf(mt[k], mt[k + w]);
					const float m = warp_min_f32(mt[0]);            // :579-584
					float left = __shfl_up_sync(0xffffffffu, prev[r][K - 1], 1);
					float right = __shfl_down_sync(0xffffffffu, prev[r][0], 1);
					if (lane == 0) left = adc_nan();                // d - 1 < 0 (:608)
					if (lane == 31) right = adc_nan();

					const uint8_t c1 = c1n[r];
					uint8_t c2[K];
#pragma unroll
					for (int k = 0; k < K; k++) c2[k] = c2n[r][k];
					fetch_classes(r, x[r] + dx, y[r] + dy);         // for step s + 1
					// penalties when the D2 class equals the D1 class (both < tau or both > tau), else middle
					const bool c1lt = c1 == 0;
					const float P1e = c1 == 1 ? P1m : (c1lt ? P1f : P1s);
					const float P2e = c1 == 1 ? P2m : (c1lt ? P2f : P2s);
					const float P1ae = c1 == 1 ? P1m_a : (c1lt ? P1f_a : P1s_a);
#pragma unroll
					for (int k = 0; k < K; k++) {
						const bool eq = c2[k] == c1;
						const float P1 = eq ? P1e : P1m, P2 = eq ? P2e : P2m, P1a = eq ? P1ae : P1m_a;
						const float pm = k > 0 ? prev[r][k - 1] : left;
						const float pp = k < K - 1 ? prev[r][k + 1] : right;
						float cost = fminf(prev[r][k], m + P2);                            // :607
						cost = fminf(cost, pm + (SD == 2 ? P1a : P1));                     // :609
						cost = fminf(cost, pp + (SD == 3 ? P1a : P1));                     // :612
						val[k] = cin[k] + cost - m;                                        // :615
					}
				}
				float o[K];
#pragma unroll
				for (int k = 0; k < K; k++) {
					o[k] = (ZERO ? 0.0f : cout[k]) + val[k];                               // :569 / :616
					prev[r][k] = val[k];                                                   // :570 / :617
				}
				store_vec<K, VEC>(o, out + base[r], dbase, D);
				// refill this ring slot with step s + PF (its values are in registers by now)
				if (s + PF < nsteps) {
					issue_vec<K, VEC>(rg[r] + (u * NV) * VSZ, in + base[r] + PF * pix_step, dbase, D);
					if (!ZERO) issue_vec<K, VEC>(rg[r] + (u * NV + 1) * VSZ, out + base[r] + PF * pix_step, dbase, D);
				}
				base[r] += pix_step;
				x[r] += dx;
				y[r] += dy;
			}
			asm volatile("cp.async.commit_group;");
		}
	}
	asm volatile("cp.async.wait_group 0;");
}

template <int K, bool VEC, int SD, bool ZERO>
int launch_pass(const uint8_t *tab, const float *in, float *out, int H, int W, int D, int pad,
		const SgmParams &prm, cudaStream_t s)
{
	constexpr int PF = K >= 16 ? 6 : 8;
	// horizontal scans have few, long lines: one warp per CTA spreads them over all SMs
	constexpr int WPB = SD < 2 ? 1 : 4;
	constexpr int NR = 1;  // NR = 2 measured 2x slower: the two recurrences are not interleaved by ptxas
	constexpr int NV = ZERO ? 1 : 2;
	constexpr int SMEM = WPB * NR * PF * NV * 32 * K * 4;
	auto kern = sgm_pass_kernel<K, VEC, SD, ZERO, PF, WPB, NR>;
	if (SMEM > 48 * 1024) {
		static bool done[64] = {false};
		int dev = 0;
		cudaGetDevice(&dev);
		if (!done[dev & 63]) {
			ADC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
			done[dev & 63] = true;
		}
	}
	const int nlines = SD < 2 ? H : W;
	kern<<<adc_div_up(nlines, WPB * NR), 32 * WPB, SMEM, s>>>(tab, in, out, H, W, D, pad, prm);
	ADC_CHECK_LAUNCH();
	return 0;
}

// pass_mask bit i = run scan direction i (0 right, 1 left, 2 down, 3 up), always in that order
template <int K, bool VEC>
int launch_all(const uint8_t *tab, const float *in, float *out, int H, int W, int D, int pad,
	       const SgmParams &prm, bool zero_out, int pass_mask, cudaStream_t s)
{
	int rc = 0;
	if (pass_mask & 1)
		rc = zero_out ? launch_pass<K, VEC, 0, true>(tab, in, out, H, W, D, pad, prm, s)
			      : launch_pass<K, VEC, 0, false>(tab, in, out, H, W, D, pad, prm, s);
	if (rc) return rc;
	if ((pass_mask & 2) && (rc = launch_pass<K, VEC, 1, false>(tab, in, out, H, W, D, pad, prm, s))) return rc;
	if ((pass_mask & 4) && (rc = launch_pass<K, VEC, 2, false>(tab, in, out, H, W, D, pad, prm, s))) return rc;
	if (pass_mask & 8) rc = launch_pass<K, VEC, 3, false>(tab, in, out, H, W, D, pad, prm, s);
	return rc;
}

}  // namespace

// lanes carry 32*K disparity slots (K from D); the class tables are padded by that many columns so
// that the padding slots d >= D also index inside the row
static int sgm_slots(int D) { return D <= 32 ? 32 : (D <= 64 ? 64 : (D <= 128 ? 128 : (D <= 256 ? 256 : 512))); }

size_t adc_sgm_table_bytes(int H, int W, int D) { return 4 * (size_t)H * (W + 2 * (size_t)sgm_slots(D)) + 16; }

int adc_sgm_classes(const float *x0, const float *x1, uint8_t *tab, int Ht, int Wt, int D, float tau_so, cudaStream_t s)
{
	const int pad = sgm_slots(D);
	const long plane = (long)Ht * (Wt + 2 * pad);
	sgm_class_kernel<<<adc_div_up(plane, 256), 256, 0, s>>>(x0, x1, tab, Ht, Wt, pad, tau_so);
	ADC_CHECK_LAUNCH();
	return 0;
}

// The selected passes over a volume (H,W,D) that is the band [yoff, yoff+H) x [xoff, xoff+W) of the
// Ht x Wt image whose class tables are in `tab` (adc_sgm_classes).  zero_out: `output` is known
// to be all zeros (main.lua:1014) -> the first (rightward) pass skips reading it.
int adc_sgm2_band(const float *in, float *out, const uint8_t *tab, int H, int W, int D, int Ht, int Wt, int yoff, int xoff,
		  float pi1, float pi2, float tau_so, float alpha1, float q1, float q2, int direction,
		  bool zero_out, int pass_mask, cudaStream_t s)
{
	SgmParams prm{pi1, pi2, tau_so, alpha1, q1, q2, direction, Ht, Wt, yoff, xoff};
	const int pad = sgm_slots(D);
	const bool vec = (D % 4 == 0) && (((uintptr_t)in | (uintptr_t)out) % 16 == 0);
	if (D <= 32) return launch_all<1, false>(tab, in, out, H, W, D, pad, prm, zero_out, pass_mask, s);
	if (D <= 64) return launch_all<2, false>(tab, in, out, H, W, D, pad, prm, zero_out, pass_mask, s);
	if (D <= 128) return vec ? launch_all<4, true>(tab, in, out, H, W, D, pad, prm, zero_out, pass_mask, s)
				 : launch_all<4, false>(tab, in, out, H, W, D, pad, prm, zero_out, pass_mask, s);
	if (D <= 256) return vec ? launch_all<8, true>(tab, in, out, H, W, D, pad, prm, zero_out, pass_mask, s)
				 : launch_all<8, false>(tab, in, out, H, W, D, pad, prm, zero_out, pass_mask, s);
	return vec ? launch_all<16, true>(tab, in, out, H, W, D, pad, prm, zero_out, pass_mask, s)
		   : launch_all<16, false>(tab, in, out, H, W, D, pad, prm, zero_out, pass_mask, s);
}

// whole image, all four directions (what adcensus.sgm2 does).  tab: scratch of adc_sgm_table_bytes(H, W, D) bytes.
int adc_sgm2(const float *x0, const float *x1, const float *in, float *out, uint8_t *tab, int H, int W, int D,
	     float pi1, float pi2, float tau_so, float alpha1, float q1, float q2, int direction,
	     bool zero_out, cudaStream_t s)
{
	int rc = adc_sgm_classes(x0, x1, tab, H, W, D, tau_so, s);
	if (rc) return rc;
	return adc_sgm2_band(in, out, tab, H, W, D, H, W, 0, 0, pi1, pi2, tau_so, alpha1, q1, q2, direction, zero_out, 15, s);
}

// ---- band-wise entry point for the row-band / column-band multi-GPU split (rowband.py) ----------
// x0, x1: the FULL Ht x Wt images (replicated on every GPU); input/output: this GPU's band volume
// (H,W,D) at image offset (yoff, xoff); pass_mask selects scan directions (bit 0 right, 1 left,
// 2 down, 3 up).  A band must contain whole scanlines of every selected pass: row bands
// (W == Wt, xoff == 0) for the horizontal passes, column bands (H == Ht, yoff == 0) for the vertical.
extern "C" int mccnn_sgm2_band(const float *x0, const float *x1, const float *input, float *output,
			       int H, int W, int D, int Ht, int Wt, int yoff, int xoff,
			       float pi1, float pi2, float tau_so, float alpha1, float sgm_q1, float sgm_q2,
			       int direction, int pass_mask, int zero_out, adcensus_stream_t stream)
{
	if (!x0 || !x1 || !input || !output || input == output) return ADCENSUS_EINVAL;
	if (H < 1 || W < 1 || D < 1 || Ht < H || Wt < W || yoff < 0 || xoff < 0 || yoff + H > Ht || xoff + W > Wt) return ADCENSUS_EINVAL;
	if ((direction != 1 && direction != -1) || (pass_mask & ~15)) return ADCENSUS_EINVAL;
	if ((pass_mask & 3) && (W != Wt || xoff != 0)) return ADCENSUS_EINVAL;   // horizontal scanlines must be whole
	if ((pass_mask & 12) && (H != Ht || yoff != 0)) return ADCENSUS_EINVAL;  // vertical scanlines must be whole
	if (D > ADCENSUS_MAX_DISP) return ADCENSUS_ELIMIT;
	cudaStream_t s = adc_stream(stream);
	uint8_t *tab = nullptr;
	int rc = adc_scratch_alloc((void **)&tab, adc_sgm_table_bytes(Ht, Wt, D), s);
	if (rc) return rc;
	rc = adc_sgm_classes(x0, x1, tab, Ht, Wt, D, tau_so, s);
	if (!rc) rc = adc_sgm2_band(input, output, tab, H, W, D, Ht, Wt, yoff, xoff, pi1, pi2, tau_so, alpha1, sgm_q1, sgm_q2,
				    direction, zero_out != 0, pass_mask, s);
	int rc2 = adc_scratch_free(tab, s);
	return rc ? rc : rc2;
}

extern "C" int adcensus_sgm2(const float *x0, const float *x1, const float *input, float *output, float *tmp,
			     int H, int W, int D, float pi1, float pi2, float tau_so, float alpha1,
			     float sgm_q1, float sgm_q2, int direction, adcensus_stream_t stream)
{
	(void)tmp;  // the reference's global line-state scratch; state lives in registers here
	if (!x0 || !x1 || !input || !output || input == output) return ADCENSUS_EINVAL;
	if (H < 1 || W < 1 || D < 1 || (direction != 1 && direction != -1)) return ADCENSUS_EINVAL;
	if (D > ADCENSUS_MAX_DISP) return ADCENSUS_ELIMIT;
	cudaStream_t s = adc_stream(stream);
	uint8_t *tab = nullptr;
	int rc = adc_scratch_alloc((void **)&tab, adc_sgm_table_bytes(H, W, D), s);
	if (rc) return rc;
	rc = adc_sgm2(x0, x1, input, output, tab, H, W, D, pi1, pi2, tau_so, alpha1, sgm_q1, sgm_q2, direction, false, s);
	int rc2 = adc_scratch_free(tab, s);
	return rc ? rc : rc2;
}

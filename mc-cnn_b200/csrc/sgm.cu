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

// State of one warp scanning one line.  SD: 0 right, 1 left, 2 down, 3 up (adcensus.cu:541-565).
template <int K, bool VEC, int SD, int PF>
struct SgmScan {
	static constexpr int VSZ = 32 * K;             // floats per cost vector slot (padded to 32*K)
	static constexpr int dx = SD == 0 ? 1 : (SD == 1 ? -1 : 0);
	static constexpr int dy = SD == 2 ? 1 : (SD == 3 ? -1 : 0);
	static constexpr int tshift_x = dx < 0 ? 1 : 0;
	static constexpr int tshift_y = dy < 0 ? 1 : 0;
	// How a lane gets the D2 classes of its K slots (bytes at consecutive table columns):
	//  SLIDE  horizontal scans: the K bytes of step s+1 are those of step s moved by one slot, so a
	//         register window is shifted and ONE new byte is loaded per step, three steps ahead;
	//  WORDS  vertical scans: K new bytes per step, fetched one step ahead as K/4+1 aligned words
	//         and funnel-shifted into place when consumed;
	//  else   (K < 4) one byte load per slot, one step ahead.
	// Nothing touches a loaded value in the step that issues the load (the table row changes every
	// step of a vertical scan, so these loads miss L1; consumed just in time they were the top stall).
	static constexpr bool SLIDE = SD < 2 && K % 4 == 0;
	static constexpr bool WORDS = SD >= 2 && K % 4 == 0;
	static constexpr int NWD = (K + 3) / 4;        // class window registers, 4 slots each
	static constexpr int NRAW = K / 4 + 1;

	const float *in;
	float *out;
	float *ring;                                   // this lane's K floats of slot 0: [PF][2][VSZ] (in, out)
	const uint8_t *t1, *t2;                        // D1 / D2 class planes, at image column xoff
	int lane, dbase, D, Wp, Ht, W, yoff, ddir;
	long pix_step;
	float P1f, P2f, P1s, P2s, P1m, P2m, P1f_a, P1s_a, P1m_a;
	int x, y;                                      // pixel of the NEXT step to execute
	long base;
	float prev[K];
	unsigned cw[NWD], c1c;                         // classes of the step being executed; byte k%4 of cw[k/4] = slot k
	unsigned nw[NWD], q1, q2, c1q[3];              // SLIDE: next window, entering bytes, D1 classes ahead
	const uint8_t *pe, *p1;
	int xl;
	bool up;
	unsigned raw[NRAW], cb[K], c1n;                // WORDS / byte path: classes of the next step
	int rsh;

	__device__ __forceinline__ void init(const uint8_t *tab, const float *in_, float *out_, float *ring_, int H, int W_, int D_,
					      int pad, const SgmParams &prm, int line)
	{
		in = in_; out = out_; ring = ring_;
		lane = threadIdx.x & 31;
		dbase = lane * K;
		D = D_;
		W = W_;
		// adcensus.cu:595-605 and :609/:612, same expressions
		P1f = prm.pi1; P2f = prm.pi2;
		P1s = prm.pi1 / (prm.q1 * prm.q2); P2s = prm.pi2 / (prm.q1 * prm.q2);
		P1m = prm.pi1 / prm.q1; P2m = prm.pi2 / prm.q1;
		P1f_a = P1f / prm.alpha1; P1s_a = P1s / prm.alpha1; P1m_a = P1m / prm.alpha1;
		// class tables: D1 from plane (SD<2 ? 0 : 1), D2 from plane (SD<2 ? 2 : 3).  The stored
		// difference at (y, j) pairs pixel j with its left / upper neighbour, so scans that look
		// right / down (dx = -1, dy = -1) read the entry one further.
		Wp = prm.Wt + 2 * pad;
		Ht = prm.Ht; yoff = prm.yoff; ddir = prm.direction;
		const long plane = (long)prm.Ht * Wp;
		t1 = tab + (SD < 2 ? 0 : 1) * plane + pad + prm.xoff;
		t2 = tab + (SD < 2 ? 2 : 3) * plane + pad + prm.xoff;
		x = SD == 0 ? 0 : (SD == 1 ? W - 1 : line);
		y = SD == 2 ? 0 : (SD == 3 ? H - 1 : line);
		pix_step = (long)(dy * W + dx) * D;
		base = ((long)y * W + x) * D + dbase;
		// NaN in the padding slots of the ring (never overwritten)
#pragma unroll
		for (int s = 0; s < PF * 2; s++)
#pragma unroll
			for (int k = 0; k < K; k++) ring[s * VSZ + k] = adc_nan();
		cls_prime(x + dx, y + dy);                     // step 0 uses no penalties
	}

	__device__ __forceinline__ int clampx(int xs) const { return min(max(xs, 0), W - 1); }

	// D1 class (:587) and D2 classes (:588-594) of pixel (xs, ys), the first recurrence step
	__device__ __forceinline__ void cls_prime(int xs, int ys)
	{
		if constexpr (SLIDE) {
			const int ty = min(max(ys + yoff, 0), Ht - 1);
			const uint8_t *r2 = t2 + (long)ty * Wp + tshift_x;
			p1 = t1 + (long)ty * Wp + tshift_x;
			up = dx * ddir < 0;                        // slot k of the next step = slot k-1 of this one
			pe = r2 + (dbase + (up ? 0 : K - 1)) * ddir;
#pragma unroll
			for (int i = 0; i < NWD; i++) nw[i] = 0;
#pragma unroll
			for (int k = 0; k < K; k++)
				nw[k / 4] |= (unsigned)__ldg(r2 + clampx(xs) + (dbase + k) * ddir) << (8 * (k & 3));
			c1q[0] = __ldg(p1 + clampx(xs));
			c1q[1] = __ldg(p1 + clampx(xs + dx));
			c1q[2] = __ldg(p1 + clampx(xs + 2 * dx));
			q1 = __ldg(pe + clampx(xs + dx));
			q2 = __ldg(pe + clampx(xs + 2 * dx));
			xl = xs + 3 * dx;
		} else {
			cls_fetch(xs, ys);
		}
	}

	__device__ __forceinline__ void cls_fetch(int xs, int ys)
	{
		const int ty = min(max(ys + yoff + tshift_y, 0), Ht - 1);           // image row of the stored difference
		const uint8_t *q = t2 + (long)ty * Wp + xs + tshift_x + dbase * ddir;
		c1n = __ldg(t1 + (long)ty * Wp + xs + tshift_x);
		if constexpr (WORDS) {
			const uintptr_t lo = (uintptr_t)(ddir > 0 ? q : q - (K - 1));   // lowest address of the K bytes
			rsh = (int)(lo & 3) * 8;
			const unsigned *al = reinterpret_cast<const unsigned *>(lo & ~(uintptr_t)3);
#pragma unroll
			for (int i = 0; i < NRAW; i++) raw[i] = __ldg(al + i);
		} else {
#pragma unroll
			for (int k = 0; k < K; k++) cb[k] = __ldg(q + k * ddir);
		}
	}

	// classes of the step about to execute -> (c1c, cw); start the loads for later steps
	__device__ __forceinline__ void cls_take()
	{
		if constexpr (SLIDE) {
#pragma unroll
			for (int i = 0; i < NWD; i++) cw[i] = nw[i];
			c1c = c1q[0];
			if (up) {
#pragma unroll
				for (int i = NWD - 1; i > 0; i--) nw[i] = __funnelshift_l(nw[i - 1], nw[i], 8);
				nw[0] = (nw[0] << 8) | q1;
			} else {
#pragma unroll
				for (int i = 0; i < NWD - 1; i++) nw[i] = __funnelshift_r(nw[i], nw[i + 1], 8);
				nw[NWD - 1] = (nw[NWD - 1] >> 8) | (q1 << 24);
			}
			q1 = q2;
			c1q[0] = c1q[1];
			c1q[1] = c1q[2];
			const int xc = clampx(xl);
			q2 = __ldg(pe + xc);
			c1q[2] = __ldg(p1 + xc);
			xl += dx;
		} else {
			c1c = c1n;
			if constexpr (WORDS) {
				unsigned w[K / 4];
#pragma unroll
				for (int i = 0; i < K / 4; i++) w[i] = __funnelshift_r(raw[i], raw[i + 1], rsh);
#pragma unroll
				for (int i = 0; i < K / 4; i++) cw[i] = ddir > 0 ? w[i] : __byte_perm(w[K / 4 - 1 - i], 0, 0x0123);
			} else {
				cw[0] = 0;
#pragma unroll
				for (int k = 0; k < K; k++) cw[0] |= cb[k] << (8 * k);
			}
			cls_fetch(x + dx, y + dy);                 // for the step after this one
		}
	}

	// steps [s_begin, s_end) of the scan; use_out: read-modify-write the accumulator (else it is
	// known to hold zeros).  Streams the cost (and accumulator) vectors of the next PF pixels into
	// the shared-memory ring with cp.async; every lane fetches exactly the elements it consumes.
	__device__ __forceinline__ void run(int s_begin, int s_end, bool use_out)
	{
#pragma unroll
		for (int u = 0; u < PF; u++) {
			if (s_begin + u < s_end) {
				issue_vec<K, VEC>(ring + (u * 2) * VSZ, in + base + u * pix_step, dbase, D);
				if (use_out) issue_vec<K, VEC>(ring + (u * 2 + 1) * VSZ, out + base + u * pix_step, dbase, D);
			}
			asm volatile("cp.async.commit_group;");
		}
		int slot = 0;
#pragma unroll 1
		for (int s = s_begin; s < s_end; s++) {
			float *rs = ring + slot * (2 * VSZ);
			slot = slot + 1 == PF ? 0 : slot + 1;
			asm volatile("cp.async.wait_group %0;" ::"n"(PF - 1));
			float cin[K], cout[K];
			read_slot<K>(cin, rs);
			if (use_out) read_slot<K>(cout, rs + VSZ);

			float val[K];
			if (s == 0) {                                   // adcensus.cu:567-572
#pragma unroll
				for (int k = 0; k < K; k++) val[k] = cin[k];
			} else {
				float mt[K];
#pragma unroll
				for (int k = 0; k < K; k++) mt[k] = prev[k];
#pragma unroll
				for (int w = K / 2; w > 0; w >>= 1)
#pragma unroll
					for (int k = 0; k < w; k++) mt[k] = fminf(mt[k], mt[k + w]);
				const float m = warp_min_f32(mt[0]);            // :579-584
				float left = __shfl_up_sync(0xffffffffu, prev[K - 1], 1);
				float right = __shfl_down_sync(0xffffffffu, prev[0], 1);
				if (lane == 0) left = adc_nan();                // d - 1 < 0 (:608)
				if (lane == 31) right = adc_nan();

				cls_take();
				// penalties when the D2 class equals the D1 class (both < tau or both > tau), else middle
				const unsigned c1 = c1c;
				const bool c1lt = c1 == 0;
				const float P1e = c1 == 1 ? P1m : (c1lt ? P1f : P1s);
				const float P2e = c1 == 1 ? P2m : (c1lt ? P2f : P2s);
				const float P1ae = c1 == 1 ? P1m_a : (c1lt ? P1f_a : P1s_a);
				unsigned xw[NWD];
#pragma unroll
				for (int i = 0; i < NWD; i++) xw[i] = cw[i] ^ (c1 * 0x01010101u);
#pragma unroll
				for (int k = 0; k < K; k++) {
					const bool eq = (xw[k / 4] & (0xffu << (8 * (k & 3)))) == 0;
					const float P1 = eq ? P1e : P1m, P2 = eq ? P2e : P2m, P1a = eq ? P1ae : P1m_a;
					const float pm = k > 0 ? prev[k - 1] : left;
					const float pp = k < K - 1 ? prev[k + 1] : right;
					float cost = fminf(prev[k], m + P2);                               // :607
					cost = fminf(cost, pm + (SD == 2 ? P1a : P1));                     // :609
					cost = fminf(cost, pp + (SD == 3 ? P1a : P1));                     // :612
					val[k] = cin[k] + cost - m;                                        // :615
				}
			}
			float o[K];
#pragma unroll
			for (int k = 0; k < K; k++) {
				o[k] = (use_out ? cout[k] : 0.0f) + val[k];                            // :569 / :616
				prev[k] = val[k];                                                      // :570 / :617
			}
			store_vec<K, VEC>(o, out + base, dbase, D);
			// refill this ring slot with step s + PF (its values are in registers by now)
			if (s + PF < s_end) {
				issue_vec<K, VEC>(rs, in + base + PF * pix_step, dbase, D);
				if (use_out) issue_vec<K, VEC>(rs + VSZ, out + base + PF * pix_step, dbase, D);
			}
			asm volatile("cp.async.commit_group;");
			base += pix_step;
			x += dx;
			y += dy;
		}
		asm volatile("cp.async.wait_group 0;");
	}
};

// One launch = one direction: one warp per scanline.  ZERO: output known to be 0 on entry.
template <int K, bool VEC, int SD, bool ZERO, int PF, int WPB>
__global__ void __launch_bounds__(32 * WPB)
sgm_pass_kernel(const uint8_t *__restrict__ tab, const float *__restrict__ in, float *__restrict__ out,
		int H, int W, int D, int pad, SgmParams prm)
{
	extern __shared__ __align__(16) float sgm_smem[];
	using Scan = SgmScan<K, VEC, SD, PF>;
	const int wib = threadIdx.x >> 5;
	const int line = blockIdx.x * WPB + wib;
	if (line >= (SD < 2 ? H : W)) return;          // whole warp
	Scan sc;
	sc.init(tab, in, out, sgm_smem + (size_t)wib * PF * 2 * Scan::VSZ + (threadIdx.x & 31) * K, H, W, D, pad, prm, line);
	sc.run(0, SD < 2 ? W : H, !ZERO);
}

// Both horizontal directions of one image row in ONE CTA (two warps), for an accumulator that is
// known to be zero on entry: warp 0 scans right, warp 1 scans left, concurrently (a horizontal
// scan is 1226 strictly serial steps on only 370 lines, so a single direction leaves most SM
// sub-partitions idle).  The reference's accumulation order out = (0 + right) + left is kept:
// in the half of the row a scan reaches FIRST it stores 0 + v; after a block barrier at the
// crossing point it adds its v to what the other scan stored.  For the left scan that is
// (0 + right) + left literally; for the right scan it is (0 + right) + (0 + left), and
// (0 + a) + b == (0 + b) + a holds in IEEE arithmetic (0 + a only canonicalises -0, + commutes).
template <int K, bool VEC, int PF>
__global__ void __launch_bounds__(64)
sgm_hpair_kernel(const uint8_t *__restrict__ tab, const float *__restrict__ in, float *__restrict__ out,
		 int H, int W, int D, int pad, SgmParams prm)
{
	extern __shared__ __align__(16) float sgm_smem[];
	const int line = blockIdx.x;
	const int wib = threadIdx.x >> 5;
	const int M = W / 2;                           // columns [0, M) are reached first by the right scan
	float *ring = sgm_smem + (size_t)wib * PF * 2 * 32 * K + (threadIdx.x & 31) * K;
	if (wib == 0) {
		SgmScan<K, VEC, 0, PF> sc;
		sc.init(tab, in, out, ring, H, W, D, pad, prm, line);
		sc.run(0, M, false);
		__syncthreads();
		sc.run(M, W, true);
	} else {
		SgmScan<K, VEC, 1, PF> sc;
		sc.init(tab, in, out, ring, H, W, D, pad, prm, line);
		sc.run(0, W - M, false);                   // columns W-1 .. M
		__syncthreads();
		sc.run(W - M, W, true);                    // columns M-1 .. 0
	}
}

static int sgm_pf_env()
{
	static int v = -1;
	if (v < 0) { const char *e = getenv("ADCENSUS_SGM_PF"); v = e ? atoi(e) : 0; }
	return v;
}

template <int K, bool VEC, int SD, bool ZERO, int PFX = 0>
int launch_pass(const uint8_t *tab, const float *in, float *out, int H, int W, int D, int pad,
		const SgmParams &prm, cudaStream_t s)
{
	if constexpr (PFX == 0 && SD < 2 && K == 8) {
		if (sgm_pf_env() == 16) return launch_pass<K, VEC, SD, ZERO, 16>(tab, in, out, H, W, D, pad, prm, s);
		if (sgm_pf_env() == 24) return launch_pass<K, VEC, SD, ZERO, 24>(tab, in, out, H, W, D, pad, prm, s);
	}
	constexpr int PF = PFX ? PFX : (K >= 16 ? 6 : 8);
	// horizontal scans have few, long lines: one warp per CTA spreads them over all SMs
	constexpr int WPB = SD < 2 ? 1 : 4;
	constexpr int SMEM = WPB * PF * 2 * 32 * K * 4;
	auto kern = sgm_pass_kernel<K, VEC, SD, ZERO, PF, WPB>;
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
	kern<<<adc_div_up(nlines, WPB), 32 * WPB, SMEM, s>>>(tab, in, out, H, W, D, pad, prm);
	ADC_CHECK_LAUNCH();
	return 0;
}

template <int K, bool VEC, int PFX = 0>
int launch_hpair(const uint8_t *tab, const float *in, float *out, int H, int W, int D, int pad,
		 const SgmParams &prm, cudaStream_t s)
{
	if constexpr (PFX == 0 && K == 8) {
		if (sgm_pf_env() == 16) return launch_hpair<K, VEC, 16>(tab, in, out, H, W, D, pad, prm, s);
		if (sgm_pf_env() == 24) return launch_hpair<K, VEC, 24>(tab, in, out, H, W, D, pad, prm, s);
	}
	constexpr int PF = PFX ? PFX : (K >= 16 ? 6 : 8);
	constexpr int SMEM = 2 * PF * 2 * 32 * K * 4;
	auto kern = sgm_hpair_kernel<K, VEC, PF>;
	if (SMEM > 48 * 1024) {
		static bool done[64] = {false};
		int dev = 0;
		cudaGetDevice(&dev);
		if (!done[dev & 63]) {
			ADC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
			done[dev & 63] = true;
		}
	}
	kern<<<H, 64, SMEM, s>>>(tab, in, out, H, W, D, pad, prm);
	ADC_CHECK_LAUNCH();
	return 0;
}

// pass_mask bit i = run scan direction i (0 right, 1 left, 2 down, 3 up), always in that order
template <int K, bool VEC>
int launch_all(const uint8_t *tab, const float *in, float *out, int H, int W, int D, int pad,
	       const SgmParams &prm, bool zero_out, int pass_mask, cudaStream_t s)
{
	int rc = 0;
	if ((pass_mask & 3) == 3 && zero_out && W >= 2) {
		rc = launch_hpair<K, VEC>(tab, in, out, H, W, D, pad, prm, s);   // right and left concurrently
	} else {
		if (pass_mask & 1)
			rc = zero_out ? launch_pass<K, VEC, 0, true>(tab, in, out, H, W, D, pad, prm, s)
				      : launch_pass<K, VEC, 0, false>(tab, in, out, H, W, D, pad, prm, s);
		if (rc) return rc;
		if ((pass_mask & 2) && (rc = launch_pass<K, VEC, 1, false>(tab, in, out, H, W, D, pad, prm, s))) return rc;
	}
	if (rc) return rc;
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

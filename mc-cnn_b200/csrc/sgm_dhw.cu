// sgm_dhw.cu -- adcensus.sgm2 (adcensus.cu:535-697) scanning the (D, H, ld) layout of the cost volume directly.
//
// main.lua:1008-1020 permutes the volume to (H, W, D) for sgm2 and back, and divides by 4.  The fused pipeline
// keeps every volume in the (D, H, ld) layout (ld = row pitch, a multiple of 4 floats), so the four permutes
// disappear, the accumulation order of the four directions (right, left, down, up; :567-617) is unchanged, and
// the /4 of main.lua:1020 rides on the last pass's store.  Arithmetic per (pixel, d) is exactly that of sgm.cu
// (and of the reference): adds, fminf, IEEE divisions => bit-identical results.
//
// Mapping.  A lane owns K consecutive disparities (lane l: d = lK .. lK+K-1), the line state L_r(p - r, .) lives
// in registers, min over d = register tree + redux.sync.min.f32, d+-1 neighbours by one shuffle each -- as in
// sgm.cu.  What changes is how data reaches the lanes: in (D, H, ld) the K values of a pixel are K different
// rows, so every lane moves 16-byte pieces = FOUR ADJACENT COLUMNS of each of its K rows:
//   * horizontal scans (one warp per image row and direction, both directions of a row in one CTA, meeting in
//     the middle like sgm_hpair_kernel): a ring slot is one group of 4 columns = 4 consecutive steps;
//   * vertical scans: a warp owns 4 adjacent columns and runs their 4 independent recurrences interleaved (ILP
//     instead of more warps); a ring slot is one image row.
// Ring slots are laid out [k][lane][4 floats]: every cp.async / LDS.128 of a warp is 512 contiguous bytes.
// Penalty classes: horizontal scans reuse the byte tables of sgm.cu (staged class rows, sliding register
// window); vertical scans fetch per step and lane one pre-compared selector record (bit c*K+k = class of D2
// at column c, slot k equals the class of D1; plus the four D1 classes) through the same ring.
#include "common.cuh"

// from sgm.cu
int adc_sgm_classes(const float *x0, const float *x1, uint8_t *tab, int Ht, int Wt, int D, float tau_so, cudaStream_t s);

namespace {

struct SgmDParams {
	float pi1, pi2, tau_so, alpha1, q1, q2;
	int direction;
};

__device__ __forceinline__ float wmin_f32(float v)
{
	float r;
	asm("redux.sync.min.f32 %0, %1, 0xffffffff;" : "=f"(r) : "f"(v));   // NaN inputs are skipped
	return r;
}
__device__ __forceinline__ void cpa16(void *smem_dst, const void *gsrc)
{
	asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc));
}
__device__ __forceinline__ void cpa8(void *smem_dst, const void *gsrc)
{
	asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc));
}
__device__ __forceinline__ float f4get(const float4 &v, int j) { return j == 0 ? v.x : (j == 1 ? v.y : (j == 2 ? v.z : v.w)); }
__device__ __forceinline__ void f4set(float4 &v, int j, float a)
{
	if (j == 0) v.x = a;
	else if (j == 1) v.y = a;
	else if (j == 2) v.z = a;
	else v.w = a;
}

// penalties of adcensus.cu:595-605 and :609/:612, same expressions
struct SgmPen {
	float P1f, P2f, P1s, P2s, P1m, P2m, P1f_a, P1s_a, P1m_a;
	__device__ __forceinline__ void init(const SgmDParams &p)
	{
		P1f = p.pi1; P2f = p.pi2;
		P1s = p.pi1 / (p.q1 * p.q2); P2s = p.pi2 / (p.q1 * p.q2);
		P1m = p.pi1 / p.q1; P2m = p.pi2 / p.q1;
		P1f_a = P1f / p.alpha1; P1s_a = P1s / p.alpha1; P1m_a = P1m / p.alpha1;
	}
};

// one step of one scanline for the K slots of a lane: prev -> val (adcensus.cu:576-617 without the accumulate).
// SD: 0 right, 1 left, 2 down, 3 up.  eqbits bit k: the D2 class of slot k equals the D1 class c1.
template <int K, int SD>
__device__ __forceinline__ void sgm_cell(const SgmPen &pn, float (&prev)[K], const float (&cin)[K], unsigned c1, unsigned eqbits, int lane)
{
	float mt[K];
#pragma unroll
	for (int k = 0; k < K; k++) mt[k] = prev[k];
#pragma unroll
	for (int w = K / 2; w > 0; w >>= 1)
#pragma unroll
		for (int k = 0; k < w; k++) mt[k] = fminf(mt[k], mt[k + w]);
	const float m = wmin_f32(mt[0]);                                   // :579-584
	float left = __shfl_up_sync(0xffffffffu, prev[K - 1], 1);
	float right = __shfl_down_sync(0xffffffffu, prev[0], 1);
	if (lane == 0) left = adc_nan();                                    // d - 1 < 0 (:608)
	if (lane == 31) right = adc_nan();
	const bool c1lt = c1 == 0;
	const float P1e = c1 == 1 ? pn.P1m : (c1lt ? pn.P1f : pn.P1s);
	const float P2e = c1 == 1 ? pn.P2m : (c1lt ? pn.P2f : pn.P2s);
	const float P1ae = c1 == 1 ? pn.P1m_a : (c1lt ? pn.P1f_a : pn.P1s_a);
	float val[K];
#pragma unroll
	for (int k = 0; k < K; k++) {
		const bool eq = (eqbits >> k) & 1u;
		const float P1 = eq ? P1e : pn.P1m, P2 = eq ? P2e : pn.P2m, P1a = eq ? P1ae : pn.P1m_a;
		const float pm = k > 0 ? prev[k - 1] : left;
		const float pp = k < K - 1 ? prev[k + 1] : right;
		float cost = fminf(prev[k], m + P2);                              // :607
		cost = fminf(cost, pm + (SD == 2 ? P1a : P1));                    // :609
		cost = fminf(cost, pp + (SD == 3 ? P1a : P1));                    // :612
		val[k] = cin[k] + cost - m;                                       // :615
	}
#pragma unroll
	for (int k = 0; k < K; k++) prev[k] = val[k];
}

// ---------------------------------------------------------------- horizontal scans
// class bytes of one image row, sliding register window (see sgm.cu SgmScan)
template <int K>
struct HCls {
	static constexpr int NWD = (K + 3) / 4;
	const uint8_t *p1, *pe;
	unsigned nw[NWD], qn, c1n;
	int xl, W, dx;
	bool up;
	__device__ __forceinline__ int clampx(int xs) const { return min(max(xs, 0), W - 1); }
	// srow1 / srow2: shared-memory D1 / D2 class rows at image column 0; x0: first pixel of the scan
	__device__ __forceinline__ void init(const uint8_t *srow1, const uint8_t *srow2, int W_, int dx_, int ddir, int dbase, int x0)
	{
		W = W_; dx = dx_;
		const int tshift = dx < 0 ? 1 : 0;             // the stored difference pairs a pixel with its LEFT neighbour
		const int xs = clampx(x0 + dx);
		p1 = srow1 + tshift;
		const uint8_t *r2 = srow2 + tshift;
		up = dx * ddir < 0;                            // slot k of the next step = slot k-1 of this one
		pe = r2 + (dbase + (up ? 0 : K - 1)) * ddir;
#pragma unroll
		for (int i = 0; i < NWD; i++) nw[i] = 0;
#pragma unroll
		for (int k = 0; k < K; k++) nw[k / 4] |= (unsigned)r2[xs + (dbase + k) * ddir] << (8 * (k & 3));   // :588-594
		c1n = p1[xs];                                  // :587
		qn = pe[clampx(xs + dx)];
		xl = xs + dx;
	}
	// classes of the step about to execute: c1 and the equality bits of the K slots; then slide the window
	__device__ __forceinline__ void take(unsigned &c1, unsigned &eqbits)
	{
		c1 = c1n;
		unsigned e = 0;
#pragma unroll
		for (int k = 0; k < K; k++) {
			const unsigned b = (nw[k / 4] >> (8 * (k & 3))) & 0xffu;
			e |= (unsigned)(b == c1) << k;
		}
		eqbits = e;
		constexpr unsigned topmask = (1u << (8 * ((K - 1) & 3))) - 1u;
		if (up) {
#pragma unroll
			for (int i = NWD - 1; i > 0; i--) nw[i] = __funnelshift_l(nw[i - 1], nw[i], 8);
			nw[0] = (nw[0] << 8) | qn;
		} else {
#pragma unroll
			for (int i = 0; i < NWD - 1; i++) nw[i] = __funnelshift_r(nw[i], nw[i + 1], 8);
			nw[NWD - 1] = ((nw[NWD - 1] >> 8) & topmask) | (qn << (8 * ((K - 1) & 3)));
		}
		c1n = p1[clampx(xl)];
		qn = pe[clampx(xl + dx)];
		xl += dx;
	}
};

// One warp, one image row, one direction (SD 0 right / 1 left).  Ring slot = one group of 4 columns:
// [2 arrays (cost, accumulator)][K][32 lanes] float4.
template <int K, int SD, int PF>
struct HScan {
	static constexpr int dx = SD == 0 ? 1 : -1;
	static constexpr int SLOT = 2 * K * 32;            // float4 per slot
	const float *in;
	float *out;
	float4 *ring;                                      // this lane's float4 of slot 0, array 0, k = 0
	long rowoff, kstride;                              // element offset of (d = dbase, y, x = 0); elements between d and d+1
	int lane, dbase, D, W;
	bool first;
	float prev[K];
	SgmPen pn;
	HCls<K> cls;

	__device__ __forceinline__ void init(const float *in_, float *out_, float4 *ring_, const uint8_t *s1, const uint8_t *s2,
					      int H, int W_, int ld, int D_, const SgmDParams &prm, int y)
	{
		in = in_; out = out_;
		lane = threadIdx.x & 31;
		ring = ring_ + lane;
		dbase = lane * K;
		D = D_; W = W_;
		kstride = (long)H * ld;
		rowoff = ((long)dbase * H + y) * ld;
		first = true;
		pn.init(prm);
		cls.init(s1, s2, W_, dx, prm.direction, dbase, SD == 0 ? 0 : W_ - 1);
		const float q = adc_nan();                      // padding slots d >= D are never loaded: they stay NaN
#pragma unroll
		for (int s = 0; s < PF; s++)
#pragma unroll
			for (int k = 0; k < 2 * K; k++) ring[s * SLOT + k * 32] = make_float4(q, q, q, q);
	}

	__device__ __forceinline__ void issue(int slot, int g, bool use_out)
	{
#pragma unroll
		for (int k = 0; k < K; k++)
			if (dbase + k < D) {
				const long e = rowoff + k * kstride + 4 * g;
				cpa16(ring + slot * SLOT + k * 32, in + e);
				if (use_out) cpa16(ring + slot * SLOT + (K + k) * 32, out + e);
			}
	}

	// groups g0, g0 + dx, ... (ng of them); only columns in [0, W) are real steps
	__device__ __forceinline__ void run(int g0, int ng, bool use_out)
	{
#pragma unroll
		for (int u = 0; u < PF; u++) {
			if (u < ng) issue(u, g0 + u * dx, use_out);
			asm volatile("cp.async.commit_group;");
		}
		int slot = 0;
#pragma unroll 1
		for (int gi = 0; gi < ng; gi++) {
			const int g = g0 + gi * dx;
			const int cs = slot;
			slot = slot + 1 == PF ? 0 : slot + 1;
			asm volatile("cp.async.wait_group %0;" ::"n"(PF - 1));
			float4 c4[K], o4[K];
#pragma unroll
			for (int k = 0; k < K; k++) c4[k] = ring[cs * SLOT + k * 32];
			if (use_out) {
#pragma unroll
				for (int k = 0; k < K; k++) o4[k] = ring[cs * SLOT + (K + k) * 32];
			} else {
#pragma unroll
				for (int k = 0; k < K; k++) o4[k] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
			}
#pragma unroll
			for (int jj = 0; jj < 4; jj++) {
				const int j = dx > 0 ? jj : 3 - jj;
				if (4 * g + j < W) {                        // warp-uniform; false only in the padding of the last group
					float cin[K];
#pragma unroll
					for (int k = 0; k < K; k++) cin[k] = f4get(c4[k], j);
					if (first) {                            // adcensus.cu:567-572
#pragma unroll
						for (int k = 0; k < K; k++) prev[k] = cin[k];
						first = false;
					} else {
						unsigned c1, eq;
						cls.take(c1, eq);
						sgm_cell<K, SD>(pn, prev, cin, c1, eq, lane);
					}
#pragma unroll
					for (int k = 0; k < K; k++) f4set(o4[k], j, f4get(o4[k], j) + prev[k]);   // :569 / :616
				}
			}
#pragma unroll
			for (int k = 0; k < K; k++)
				if (dbase + k < D) *reinterpret_cast<float4 *>(out + rowoff + k * kstride + 4 * g) = o4[k];
			if (gi + PF < ng) issue(cs, g + PF * dx, use_out);
			asm volatile("cp.async.commit_group;");
		}
		asm volatile("cp.async.wait_group 0;");
	}
};

__host__ __device__ inline int sgmd_row_bytes(int W, int pad) { return (W + 2 * pad + 8 + 15) & ~15; }

__device__ __forceinline__ const uint8_t *sgmd_stage_row(unsigned *dst, const uint8_t *row, int nbytes, int tid, int nthreads)
{
	const uintptr_t a = (uintptr_t)row;
	const unsigned *al = reinterpret_cast<const unsigned *>(a & ~(uintptr_t)3);
	const int mis = (int)(a & 3);
	const int nw = (mis + nbytes + 3) >> 2;
	for (int i = tid; i < nw; i += nthreads) dst[i] = __ldg(al + i);
	return reinterpret_cast<const uint8_t *>(dst) + mis;
}

// Both horizontal directions of one image row in ONE CTA (two warps), accumulator known to be zero on entry
// (main.lua:1014): in the half of the row a scan reaches FIRST it stores 0 + v; after a block barrier at the
// crossing point it adds its v to what the other scan stored.  (0 + right) + left for the left scan literally,
// (0 + right) + (0 + left) for the right scan, equal because (0 + a) + b == (0 + b) + a in IEEE arithmetic.
template <int K, int PF>
__global__ void __launch_bounds__(64)
sgmd_hpair_kernel(const uint8_t *__restrict__ tab, const float *__restrict__ in, float *__restrict__ out,
		  int H, int W, int ld, int D, int pad, SgmDParams prm)
{
	extern __shared__ __align__(16) float4 sgmd_smem[];
	const int y = blockIdx.x;
	const int wib = threadIdx.x >> 5;
	constexpr int SLOT = 2 * K * 32;
	float4 *ring = sgmd_smem + (size_t)wib * PF * SLOT;
	const int Wp = W + 2 * pad, RB = sgmd_row_bytes(W, pad);
	const long plane = (long)H * Wp;
	unsigned *rows = reinterpret_cast<unsigned *>(sgmd_smem + (size_t)2 * PF * SLOT);
	const uint8_t *s1 = sgmd_stage_row(rows, tab + (long)y * Wp, Wp, threadIdx.x, 64) + pad;                      // plane 0: D1
	const uint8_t *s2 = sgmd_stage_row(rows + RB / 4, tab + 2 * plane + (long)y * Wp, Wp, threadIdx.x, 64) + pad;   // plane 2: D2
	__syncthreads();
	const int NG = (W + 3) >> 2;                       // column groups
	const int MG = NG / 2;                             // groups [0, MG) are reached first by the right scan
	if (wib == 0) {
		HScan<K, 0, PF> sc;
		sc.init(in, out, ring, s1, s2, H, W, ld, D, prm, y);
		sc.run(0, MG, false);
		__syncthreads();
		sc.run(MG, NG - MG, true);
	} else {
		HScan<K, 1, PF> sc;
		sc.init(in, out, ring, s1, s2, H, W, ld, D, prm, y);
		sc.run(NG - 1, NG - MG, false);                // groups NG-1 .. MG
		__syncthreads();
		sc.run(MG - 1, MG, true);                      // groups MG-1 .. 0
	}
}

// ---------------------------------------------------------------- vertical scans
// selector records: per (table row ty, column group xg, lane) SW words: bit c*K + k of the first words = the D2
// class of (column 4xg + c, slot k) equals the D1 class of that column; the word after them holds the four D1
// classes, one byte each.
template <int K>
struct VSel {
	static constexpr int NBW = (4 * K + 31) / 32;      // words of equality bits
	static constexpr int SW = NBW == 1 ? 2 : 4;        // record size in words (8 or 16 bytes)
};

template <int K>
__global__ void sgmd_sel_kernel(const uint8_t *__restrict__ tab, unsigned *__restrict__ sel, int H, int W, int Wp, int pad, int ddir)
{
	constexpr int NBW = VSel<K>::NBW, SW = VSel<K>::SW;
	const int lane = threadIdx.x, xg = blockIdx.x * 8 + threadIdx.y, ty = blockIdx.y;
	const int NG = (W + 3) >> 2;
	if (xg >= NG) return;
	const long plane = (long)H * Wp;
	unsigned w[SW];
#pragma unroll
	for (int i = 0; i < SW; i++) w[i] = 0;
#pragma unroll
	for (int c = 0; c < 4; c++) {
		const int x = 4 * xg + c;
		if (x >= W) continue;
		const long col = (long)ty * Wp + pad + x;
		const unsigned c1 = __ldg(tab + plane + col);                    // v0: D1 class (:587)
		const uint8_t *q = tab + 3 * plane + col + (long)lane * K * ddir;   // v1: D2 classes (:588-594), slot k at q[k * ddir]
#pragma unroll
		for (int k = 0; k < K; k++) {
			const int b = c * K + k;
			w[b >> 5] |= (unsigned)(__ldg(q + k * ddir) == c1) << (b & 31);
		}
		w[NBW] |= c1 << (8 * c);
	}
	unsigned *dst = sel + (((long)ty * NG + xg) * 32 + lane) * SW;
#pragma unroll
	for (int i = 0; i < SW; i++) dst[i] = w[i];
}

// One warp = 4 adjacent columns, all rows, one direction (SD 2 down / 3 up).  Ring slot = one image row:
// [2 arrays][K][32 lanes] float4 (the 4 columns) + [32 lanes] selector record.  LAST: multiply the final sum by
// 0.25 (exactly x / 4, main.lua:1020).
template <int K, int SD, int PF, int WPB, bool LAST>
__global__ void __launch_bounds__(32 * WPB)
sgmd_vpass_kernel(const unsigned *__restrict__ sel, const float *__restrict__ in, float *__restrict__ out,
		  int H, int W, int ld, int D, SgmDParams prm)
{
	extern __shared__ __align__(16) float4 sgmd_smem[];
	constexpr int SW = VSel<K>::SW, NBW = VSel<K>::NBW;
	constexpr int SLOT = 2 * K * 32 + 32 * SW / 4;     // float4 per slot
	constexpr int dy = SD == 2 ? 1 : -1;
	const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const int NG = (W + 3) >> 2;
	const int xg = blockIdx.x * WPB + wib;
	if (xg >= NG) return;                              // whole warp
	float4 *ring = sgmd_smem + (size_t)wib * PF * SLOT + lane;
	const int dbase = lane * K;
	const long kstride = (long)H * ld;
	SgmPen pn;
	pn.init(prm);
	{
		const float q = adc_nan();
#pragma unroll
		for (int s = 0; s < PF; s++)
#pragma unroll
			for (int k = 0; k < 2 * K; k++) ring[s * SLOT + k * 32] = make_float4(q, q, q, q);
	}
	// the stored vertical difference pairs a pixel with its UPPER neighbour: the up scan reads one row further
	constexpr int tshift = dy < 0 ? 1 : 0;
	auto issue = [&](int slot, int s) {
		const int y = SD == 2 ? s : H - 1 - s;
#pragma unroll
		for (int k = 0; k < K; k++)
			if (dbase + k < D) {
				const long e = ((long)(dbase + k) * H + y) * ld + 4 * xg;
				cpa16(ring + slot * SLOT + k * 32, in + e);
				cpa16(ring + slot * SLOT + (K + k) * 32, out + e);
			}
		if (s >= 1) {                                  // step 0 uses no penalties
			const unsigned *rec = sel + (((long)(y + tshift) * NG + xg) * 32 + lane) * SW;
			unsigned *dst = reinterpret_cast<unsigned *>(sgmd_smem + (size_t)wib * PF * SLOT + slot * SLOT + 2 * K * 32) + lane * SW;
			if (SW == 2) cpa8(dst, rec);
			else cpa16(dst, rec);
		}
	};
	float prev[4][K];
#pragma unroll
	for (int u = 0; u < PF; u++) {
		if (u < H) issue(u, u);
		asm volatile("cp.async.commit_group;");
	}
	int slot = 0;
#pragma unroll 1
	for (int s = 0; s < H; s++) {
		const int y = SD == 2 ? s : H - 1 - s;
		const int cs = slot;
		slot = slot + 1 == PF ? 0 : slot + 1;
		asm volatile("cp.async.wait_group %0;" ::"n"(PF - 1));
		float4 c4[K], o4[K];
#pragma unroll
		for (int k = 0; k < K; k++) {
			c4[k] = ring[cs * SLOT + k * 32];
			o4[k] = ring[cs * SLOT + (K + k) * 32];
		}
		if (s == 0) {                                   // adcensus.cu:567-572
#pragma unroll
			for (int c = 0; c < 4; c++)
#pragma unroll
				for (int k = 0; k < K; k++) prev[c][k] = f4get(c4[k], c);
		} else {
			unsigned w[SW];
			const unsigned *rec = reinterpret_cast<const unsigned *>(sgmd_smem + (size_t)wib * PF * SLOT + cs * SLOT + 2 * K * 32) + lane * SW;
#pragma unroll
			for (int i = 0; i < SW; i++) w[i] = rec[i];
#pragma unroll
			for (int c = 0; c < 4; c++) {
				float cin[K];
#pragma unroll
				for (int k = 0; k < K; k++) cin[k] = f4get(c4[k], c);
				const unsigned c1 = (w[NBW] >> (8 * c)) & 3u;
				unsigned eq;
				if (K >= 8) eq = (w[(c * K) >> 5] >> ((c * K) & 31));
				else eq = w[0] >> (c * K);
				sgm_cell<K, SD>(pn, prev[c], cin, c1, eq, lane);
			}
		}
#pragma unroll
		for (int k = 0; k < K; k++) {
			float4 o = o4[k];
			o.x += prev[0][k]; o.y += prev[1][k]; o.z += prev[2][k]; o.w += prev[3][k];   // :569 / :616
			if (LAST) { o.x *= 0.25f; o.y *= 0.25f; o.z *= 0.25f; o.w *= 0.25f; }              // main.lua:1020, exact
			if (dbase + k < D) *reinterpret_cast<float4 *>(out + ((long)(dbase + k) * H + y) * ld + 4 * xg) = o;
		}
		if (s + PF < H) issue(cs, s + PF);
		asm volatile("cp.async.commit_group;");
	}
	asm volatile("cp.async.wait_group 0;");
}

constexpr int SGMD_SMEM_MAX = 200 * 1024;

template <typename Kern>
int sgmd_allow_smem(Kern kern, int smem, bool *done)
{
	if (smem > SGMD_SMEM_MAX) return ADCENSUS_ELIMIT;
	int dev = 0;
	cudaGetDevice(&dev);
	if (!done[dev & 63]) {
		ADC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SGMD_SMEM_MAX));
		done[dev & 63] = true;
	}
	return 0;
}

__host__ __device__ inline size_t sgmd_sel_offset(int H, int Wp) { return ((size_t)4 * H * Wp + 16 + 255) & ~(size_t)255; }

template <int K>
int launch_dhw(const uint8_t *tab, const float *in, float *acc, int H, int W, int ld, int D, int pad, const SgmDParams &prm,
	       bool div4, cudaStream_t s)
{
	constexpr int PFH = K >= 16 ? 3 : 4, PFV = K >= 16 ? 4 : 6, WPB = 2;
	constexpr int SW = VSel<K>::SW;
	const int Wp = W + 2 * pad, NG = (W + 3) >> 2;
	unsigned *sel = reinterpret_cast<unsigned *>(const_cast<uint8_t *>(tab) + sgmd_sel_offset(H, Wp));
	if (H > 65535) return ADCENSUS_ELIMIT;
	{
		const dim3 grid(adc_div_up(NG, 8), H), block(32, 8);
		sgmd_sel_kernel<K><<<grid, block, 0, s>>>(tab, sel, H, W, Wp, pad, prm.direction);
		ADC_CHECK_LAUNCH();
	}
	{
		const int smem = 2 * PFH * (2 * K * 32) * 16 + 2 * sgmd_row_bytes(W, pad);
		auto kern = sgmd_hpair_kernel<K, PFH>;
		static bool done[64] = {false};
		int rc = sgmd_allow_smem(kern, smem, done);
		if (rc) return rc;
		kern<<<H, 64, smem, s>>>(tab, in, acc, H, W, ld, D, pad, prm);
		ADC_CHECK_LAUNCH();
	}
	const int smemv = WPB * PFV * (2 * K * 32 + 32 * SW / 4) * 16;
	{
		auto kern = sgmd_vpass_kernel<K, 2, PFV, WPB, false>;
		static bool done[64] = {false};
		int rc = sgmd_allow_smem(kern, smemv, done);
		if (rc) return rc;
		kern<<<adc_div_up(NG, WPB), 32 * WPB, smemv, s>>>(sel, in, acc, H, W, ld, D, prm);
		ADC_CHECK_LAUNCH();
	}
	if (div4) {
		auto kern = sgmd_vpass_kernel<K, 3, PFV, WPB, true>;
		static bool done[64] = {false};
		int rc = sgmd_allow_smem(kern, smemv, done);
		if (rc) return rc;
		kern<<<adc_div_up(NG, WPB), 32 * WPB, smemv, s>>>(sel, in, acc, H, W, ld, D, prm);
	} else {
		auto kern = sgmd_vpass_kernel<K, 3, PFV, WPB, false>;
		static bool done[64] = {false};
		int rc = sgmd_allow_smem(kern, smemv, done);
		if (rc) return rc;
		kern<<<adc_div_up(NG, WPB), 32 * WPB, smemv, s>>>(sel, in, acc, H, W, ld, D, prm);
	}
	ADC_CHECK_LAUNCH();
	return 0;
}

int sgmd_slots(int D) { return D <= 32 ? 32 : (D <= 64 ? 64 : (D <= 128 ? 128 : (D <= 256 ? 256 : 512))); }

}  // namespace

// class planes (as sgm.cu) + selector records of the vertical scans
size_t adc_sgm_dhw_table_bytes(int H, int W, int D)
{
	const int pad = sgmd_slots(D), K = pad / 32;
	const size_t sw = (4 * K + 31) / 32 == 1 ? 2 : 4;
	return sgmd_sel_offset(H, W + 2 * pad) + (size_t)(H + 1) * ((W + 3) / 4) * 32 * sw * 4;
}

// sgm2 on (D, H, ld) volumes: acc = sum of the four directional costs of `in` (acc need not be initialised),
// times 1/4 when div4 (main.lua:1014-1020 without the permutes).  in / acc: 16-byte aligned, ld % 4 == 0, ld >= W rounded
// up to 4; the padding columns of acc receive unspecified values.  tab: adc_sgm_dhw_table_bytes(H, W, D) bytes of scratch.
int adc_sgm2_dhw(const float *x0, const float *x1, const float *in, float *acc, uint8_t *tab, int H, int W, int ld, int D,
		 float pi1, float pi2, float tau_so, float alpha1, float q1, float q2, int direction, bool div4, cudaStream_t s)
{
	if ((ld & 3) || ld < ((W + 3) & ~3) || ((((uintptr_t)in) | ((uintptr_t)acc)) & 15)) return ADCENSUS_EINVAL;
	int rc = adc_sgm_classes(x0, x1, tab, H, W, D, tau_so, s);
	if (rc) return rc;
	const SgmDParams prm{pi1, pi2, tau_so, alpha1, q1, q2, direction};
	const int pad = sgmd_slots(D);
	switch (pad / 32) {
	case 1: return launch_dhw<1>(tab, in, acc, H, W, ld, D, pad, prm, div4, s);
	case 2: return launch_dhw<2>(tab, in, acc, H, W, ld, D, pad, prm, div4, s);
	case 4: return launch_dhw<4>(tab, in, acc, H, W, ld, D, pad, prm, div4, s);
	case 8: return launch_dhw<8>(tab, in, acc, H, W, ld, D, pad, prm, div4, s);
	default: return launch_dhw<16>(tab, in, acc, H, W, ld, D, pad, prm, div4, s);
	}
}

// public: sgm2 + /4 on pitched (D, H, ld) volumes (what the fused pipeline runs between its CBCA blocks)
extern "C" int mccnn_sgm2_dhw(const float *x0, const float *x1, const float *input, float *output,
			      int H, int W, int ld, int D, float pi1, float pi2, float tau_so, float alpha1,
			      float sgm_q1, float sgm_q2, int direction, int div4, adcensus_stream_t stream)
{
	if (!x0 || !x1 || !input || !output || input == output) return ADCENSUS_EINVAL;
	if (H < 1 || W < 1 || D < 1 || (direction != 1 && direction != -1)) return ADCENSUS_EINVAL;
	if (D > ADCENSUS_MAX_DISP) return ADCENSUS_ELIMIT;
	cudaStream_t s = adc_stream(stream);
	uint8_t *tab = nullptr;
	int rc = adc_scratch_alloc((void **)&tab, adc_sgm_dhw_table_bytes(H, W, D), s);
	if (rc) return rc;
	rc = adc_sgm2_dhw(x0, x1, input, output, tab, H, W, ld, D, pi1, pi2, tau_so, alpha1, sgm_q1, sgm_q2, direction, div4 != 0, s);
	int rc2 = adc_scratch_free(tab, s);
	return rc ? rc : rc2;
}

// adcensus_cost.cu -- the net-free matching costs adcensus.ad / adcensus.census (adcensus.cu:62-175) for sm_100a.
//
// The reference walks the 9x9 window per (d, y, x) output: 81 (ad) or 81 * nch * 2 (census) global loads each.
//
// census: the comparisons I(q) < I(p) do not depend on d, so they are done ONCE per image: census_pack_kernel turns
//   every pixel's window into an 81-bit descriptor (3 words per channel, tap order = the reference's loop order) plus,
//   per pixel, the 81-bit mask of taps that lie inside the image.  A cost is then
//       taps outside either image (count as mismatches, :139)  = 81 - popc(V(y,x) & V(y,x+d))
//       + mismatching comparisons (:137)                         = popc((B0(y,x) ^ B1(y,x+d)) & V & V)
//   summed over the channels and divided by nch: integer work, bit-identical by construction (the reference counts in a
//   float, which is exact for these small integers).
// ad: the mean absolute difference over the in-image taps, accumulated in the reference's order (rows outer, columns
//   inner, one fp32 accumulator, :78-85) => bit-identical.  Per disparity the CTA first builds the tile of
//   |x0(y,x) - x1(y,x+d)| (+0.0f where a tap is outside either image: adding it leaves the non-negative accumulator
//   unchanged) in shared memory; a thread then owns 4 adjacent outputs and reads each window row once (3 LDS.128 for
//   12 values) for all four, so an output costs 81 FADD + 7 LDS instead of 162 global loads.
#include "common.cuh"

namespace {

// ------------------------------------------------------------------ census
// words: [nch][3][H][W] descriptors; vmask: [3][H][W] in-image masks (shared by both images: same H, W)
__global__ void census_pack_kernel(const float *__restrict__ img, uint32_t *__restrict__ words, uint32_t *__restrict__ vmask,
				   int nch, int H, int W)
{
	const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, ch = blockIdx.z;
	if (x >= W) return;
	const long HW = (long)H * W;
	const float *p = img + ch * HW;
	const float c = __ldg(p + (long)y * W + x);
	uint32_t b[3] = {0, 0, 0}, v[3] = {0, 0, 0};
	int t = 0;
	for (int yy = y - 4; yy <= y + 4; yy++)
		for (int xx = x - 4; xx <= x + 4; xx++, t++) {
			if (0 <= xx && xx < W && 0 <= yy && yy < H) {
				v[t >> 5] |= 1u << (t & 31);
				if (__ldg(p + (long)yy * W + xx) < c) b[t >> 5] |= 1u << (t & 31);   // :137 x0[ind_q] < x0[ind_p]
			}
		}
	const long pix = (long)y * W + x;
#pragma unroll
	for (int i = 0; i < 3; i++) {
		words[((long)ch * 3 + i) * HW + pix] = b[i];
		if (vmask && ch == 0) vmask[i * HW + pix] = v[i];
	}
}

constexpr int CENSUS_MAXCH = 4;

// one thread per pixel (y, x), looping over the disparities of its chunk (blockIdx.z)
template <int NCH>
__global__ void census_cost_kernel(const uint32_t *__restrict__ w0, const uint32_t *__restrict__ w1, const uint32_t *__restrict__ vmask,
				   float *__restrict__ out, int D, int H, int W, int direction, int dch, int nch_rt)
{
	const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
	if (x >= W) return;
	const int nch = NCH > 0 ? NCH : nch_rt;
	const long HW = (long)H * W, pix = (long)y * W + x;
	uint32_t b0[(NCH > 0 ? NCH : CENSUS_MAXCH) * 3], v0[3];
#pragma unroll
	for (int i = 0; i < 3; i++) v0[i] = __ldg(vmask + i * HW + pix);
#pragma unroll
	for (int c = 0; c < (NCH > 0 ? NCH : CENSUS_MAXCH); c++)
#pragma unroll
		for (int i = 0; i < 3; i++) b0[c * 3 + i] = c < nch ? __ldg(w0 + ((long)c * 3 + i) * HW + pix) : 0u;
	const int d0 = blockIdx.z * dch, d1 = min(D, d0 + dch);
	const float fn = (float)nch;
	for (int dd = d0; dd < d1; dd++) {
		const int xs = x + dd * direction;
		float dist;
		if (0 <= xs && xs < W) {
			const long q = (long)y * W + xs;
			uint32_t val[3];
			int cnt = 0;
#pragma unroll
			for (int i = 0; i < 3; i++) {
				val[i] = v0[i] & __ldg(vmask + i * HW + q);
				cnt += __popc(val[i]);
			}
			cnt = nch * (81 - cnt);                                   // taps outside either image: dist++ for every channel (:139)
#pragma unroll
			for (int c = 0; c < (NCH > 0 ? NCH : CENSUS_MAXCH); c++)
				if (c < nch) {
#pragma unroll
					for (int i = 0; i < 3; i++) cnt += __popc((b0[c * 3 + i] ^ __ldg(w1 + ((long)c * 3 + i) * HW + q)) & val[i]);   // :137
				}
			dist = (float)cnt / fn;                                   // :143
		} else {
			dist = adc_nan();                                         // :145
		}
		out[(long)dd * HW + pix] = dist;
	}
}

// ------------------------------------------------------------------ ad
constexpr int AD_TX = 128, AD_TY = 8, AD_NT = 256, AD_R = 4;
constexpr int AD_TW = AD_TX + 2 * AD_R;            // 136 tile columns
constexpr int AD_TH = AD_TY + 2 * AD_R;            // 16 tile rows

__global__ void __launch_bounds__(AD_NT)
ad_kernel(const float *__restrict__ x0, const float *__restrict__ x1, float *__restrict__ out, int D, int H, int W, int direction, int dch)
{
	__shared__ __align__(16) float sa[AD_TH][AD_TW];   // |x0 - x1(d)| of the tile + halo, +0.0f outside either image
	const int tid = threadIdx.x;
	const int X0 = blockIdx.x * AD_TX, Y0 = blockIdx.y * AD_TY;
	const int d0 = blockIdx.z * dch, d1 = min(D, d0 + dch);
	const int tx = tid & 31, ty = tid >> 5;            // 4 outputs x = X0 + 4 tx .. + 3, row y = Y0 + ty
	const int y = Y0 + ty;
	const long HW = (long)H * W;
	// rows of the window inside the image (:78: 0 <= yy < H)
	const int ny = min(H - 1, y + AD_R) - max(0, y - AD_R) + 1;
	for (int dd = d0; dd < d1; dd++) {
		const int d = dd * direction;
		__syncthreads();                               // previous disparity's tile consumed
		for (int i = tid; i < AD_TH * AD_TW; i += AD_NT) {
			const int r = i / AD_TW, c = i - r * AD_TW;
			const int yy = Y0 - AD_R + r, xx = X0 - AD_R + c;
			float a = 0.0f;
			if (0 <= xx && xx < W && 0 <= xx + d && xx + d < W && 0 <= yy && yy < H)   // :78
				a = fabsf(__ldg(x0 + (long)yy * W + xx) - __ldg(x1 + (long)yy * W + xx + d));   // :80
			sa[r][c] = a;
		}
		__syncthreads();
		float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
		for (int r = 0; r < 2 * AD_R + 1; r++) {           // window rows ascending (:76), columns ascending (:77)
			const float4 q0 = *reinterpret_cast<const float4 *>(&sa[ty + r][4 * tx]);
			const float4 q1 = *reinterpret_cast<const float4 *>(&sa[ty + r][4 * tx + 4]);
			const float4 q2 = *reinterpret_cast<const float4 *>(&sa[ty + r][4 * tx + 8]);
			const float w[12] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w};
#pragma unroll
			for (int c = 0; c < 2 * AD_R + 1; c++)
#pragma unroll
				for (int o = 0; o < 4; o++) acc[o] += w[o + c];      // :80, one accumulator per output
		}
		if (y < H) {
#pragma unroll
			for (int o = 0; o < 4; o++) {
				const int x = X0 + 4 * tx + o;
				if (x >= W) continue;
				float dist;
				if (0 <= x + d && x + d < W) {
					// taps with 0 <= xx < W and 0 <= xx + d < W (:78): the intersection of three intervals
					const int lo = max(max(x - AD_R, 0), -d), hi = min(min(x + AD_R, W - 1), W - 1 - d);
					const int cnt = ny * (hi - lo + 1);
					dist = acc[o] / (float)cnt;                           // :86
				} else {
					dist = adc_nan();                                     // :88
				}
				out[(long)dd * HW + (long)y * W + x] = dist;
			}
		}
	}
}

}  // namespace

extern "C" int adcensus_ad(const float *x0, const float *x1, float *out, int D, int H, int W, int direction, adcensus_stream_t stream)
{
	if (!x0 || !x1 || !out || D < 1 || H < 1 || W < 1 || (direction != 1 && direction != -1)) return ADCENSUS_EINVAL;
	const int dch = 8;
	dim3 grid(adc_div_up(W, AD_TX), adc_div_up(H, AD_TY), adc_div_up(D, dch));
	if (grid.y > 65535 || grid.z > 65535) return ADCENSUS_ELIMIT;
	ad_kernel<<<grid, AD_NT, 0, adc_stream(stream)>>>(x0, x1, out, D, H, W, direction, dch);
	ADC_CHECK_LAUNCH();
	return 0;
}

extern "C" int adcensus_census(const float *x0, const float *x1, float *out, int D, int nch, int H, int W, int direction,
			       adcensus_stream_t stream)
{
	if (!x0 || !x1 || !out || D < 1 || nch < 1 || H < 1 || W < 1 || (direction != 1 && direction != -1)) return ADCENSUS_EINVAL;
	if (nch > CENSUS_MAXCH || H > 65535) return ADCENSUS_ELIMIT;         // main.lua feeds 1 (grey) or 3 (colour) channels
	cudaStream_t s = adc_stream(stream);
	const long HW = (long)H * W;
	uint32_t *buf = nullptr;
	const size_t words = (size_t)(2 * nch * 3 + 3) * HW;
	int rc = adc_scratch_alloc((void **)&buf, words * sizeof(uint32_t), s);
	if (rc) return rc;
	uint32_t *w0 = buf, *w1 = buf + (size_t)nch * 3 * HW, *vm = w1 + (size_t)nch * 3 * HW;
	dim3 pg(adc_div_up(W, 128), H, nch);
	census_pack_kernel<<<pg, 128, 0, s>>>(x0, w0, vm, nch, H, W);
	census_pack_kernel<<<pg, 128, 0, s>>>(x1, w1, nullptr, nch, H, W);
	const int dch = 16;
	dim3 grid(adc_div_up(W, 128), H, adc_div_up(D, dch));
	if (grid.z > 65535) {
		adc_scratch_free(buf, s);
		return ADCENSUS_ELIMIT;
	}
	if (nch == 1) census_cost_kernel<1><<<grid, 128, 0, s>>>(w0, w1, vm, out, D, H, W, direction, dch, nch);
	else if (nch == 3) census_cost_kernel<3><<<grid, 128, 0, s>>>(w0, w1, vm, out, D, H, W, direction, dch, nch);
	else census_cost_kernel<0><<<grid, 128, 0, s>>>(w0, w1, vm, out, D, H, W, direction, dch, nch);
	rc = (int)cudaPeekAtLastError();
	int rc2 = adc_scratch_free(buf, s);
	return rc ? rc : rc2;
}

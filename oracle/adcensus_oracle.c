/*
 * adcensus_oracle.c -- TEST INFRASTRUCTURE, not product code.
 *
 * CPU restatement of the reference's stereo-method operators.  Each function
 * cites the reference lines it follows (adcensus.cu / main.lua under
 * /root/reference).  It is the checker for the CUDA path; nothing the product
 * ships may call it (see adcensus_oracle.h).
 *
 * Numerics: the reference is built by nvcc with its default -fmad=true, so
 * `a += b * c` is one fused multiply-add on the GPU.  This file is compiled
 * with -ffp-contract=off and writes fmaf() explicitly at exactly those sites,
 * so that it reproduces the reference's fp32 results bit for bit (checked
 * against tests/golden/, produced by the reference kernels on a B200).
 * Division and sqrt are IEEE-correct on both sides (nvcc default
 * -prec-div=true -prec-sqrt=true).  `min`/`max` on floats in CUDA device code
 * are fminf/fmaxf (NaN-ignoring), `abs(float)` is fabsf, `round(float)` is
 * roundf (half away from zero).
 */
#include "adcensus_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define IDX3(a, b, c, B, C) ((((long)(a)) * (B) + (b)) * (long)(C) + (c))

/* adcensus.cu:47-60 -- ascending selection sort */
static void sel_sort(float *x, int n)
{
	for (int i = 0; i < n - 1; i++) {
		int mn = i;
		for (int j = i + 1; j < n; j++)
			if (x[j] < x[mn]) mn = j;
		float t = x[mn];
		x[mn] = x[i];
		x[i] = t;
	}
}

/* adcensus.cu:1284-1308 (Normalize_get_norm_, Normalize_forward_), host 1310-1333 */
void orc_normalize_forward(const float *in, float *norm, float *out, int N, int C, int H, int W)
{
	long HW = (long)H * W;
#pragma omp parallel for collapse(2)
	for (int n = 0; n < N; n++)
		for (long p = 0; p < HW; p++) {
			float sum = 0.0f;
			for (int c = 0; c < C; c++) {
				float x = in[((long)n * C + c) * HW + p];
				sum = fmaf(x, x, sum);            /* :1294 sum += x * x */
			}
			norm[n * HW + p] = (float)((double)sum + 1e-5); /* :1296 double literal */
		}
#pragma omp parallel for collapse(2)
	for (int n = 0; n < N; n++)
		for (long i = 0; i < (long)C * HW; i++) {
			long p = i % HW;
			out[(long)n * C * HW + i] = in[(long)n * C * HW + i] / sqrtf(norm[n * HW + p]); /* :1306 */
		}
}

/* adcensus.cu:1455-1477 StereoJoin_.  Entries with x - d < 0 are not written
 * (the caller pre-fills NaN, main.lua:946). */
void orc_stereo_join(const float *L, const float *R, float *outL, float *outR, int C, int D, int H, int W)
{
	long HW = (long)H * W;
#pragma omp parallel for
	for (long id = 0; id < HW; id++) {
		int x = (int)(id % W);
		for (int d = 0; d < D; d++) {
			if (x - d >= 0) {
				float sum = 0;
				for (int c = 0; c < C; c++)
					sum = fmaf(-L[c * HW + id], R[c * HW + id - d], sum); /* :1470 sum -= l * r */
				outL[d * HW + id] = sum;      /* :1472 */
				outR[d * HW + id - d] = sum;  /* :1473 */
			}
		}
	}
}

/* main.lua:922-927.  direction -1: columns W-1..W-n <- column W-n-1;
 * direction +1: columns 0..n-1 <- column n. */
void orc_fix_border(float *vol, int D, int H, int W, int n, int direction)
{
	for (int i = 1; i <= n; i++) {
		int dst = direction > 0 ? i - 1 : W - i;
		int src = direction > 0 ? n : W - n - 1;
		for (long r = 0; r < (long)D * H; r++) vol[r * W + dst] = vol[r * W + src];
	}
}

/* adcensus.cu:62-93 */
void orc_ad(const float *x0, const float *x1, float *out, int D, int H, int W, int direction)
{
#pragma omp parallel for collapse(2)
	for (int dd = 0; dd < D; dd++)
		for (int y = 0; y < H; y++)
			for (int x = 0; x < W; x++) {
				int d = dd * direction;
				float dist;
				if (0 <= x + d && x + d < W) {
					int cnt = 0;
					dist = 0;
					for (int yy = y - 4; yy <= y + 4; yy++)
						for (int xx = x - 4; xx <= x + 4; xx++)
							if (0 <= xx && xx < W && 0 <= xx + d && xx + d < W && 0 <= yy && yy < H) {
								int ind = yy * W + xx;
								dist += fabsf(x0[ind] - x1[ind + d]);
								cnt++;
							}
					dist /= cnt;
				} else {
					dist = NAN;
				}
				out[IDX3(dd, y, x, H, W)] = dist;
			}
}

/* adcensus.cu:117-153 */
void orc_census(const float *x0, const float *x1, float *out, int D, int nch, int H, int W, int direction)
{
#pragma omp parallel for collapse(2)
	for (int dd = 0; dd < D; dd++)
		for (int y = 0; y < H; y++)
			for (int x = 0; x < W; x++) {
				int d = dd * direction;
				float dist;
				if (0 <= x + d && x + d < W) {
					dist = 0;
					for (int i = 0; i < nch; i++) {
						long ind_p = IDX3(i, y, x, H, W);
						for (int yy = y - 4; yy <= y + 4; yy++)
							for (int xx = x - 4; xx <= x + 4; xx++) {
								if (0 <= xx && xx < W && 0 <= xx + d && xx + d < W && 0 <= yy && yy < H) {
									long ind_q = IDX3(i, yy, xx, H, W);
									if ((x0[ind_q] < x0[ind_p]) != (x1[ind_q + d] < x1[ind_p + d])) dist++;
								} else {
									dist++;
								}
							}
					}
					dist /= nch;
				} else {
					dist = NAN;
				}
				out[IDX3(dd, y, x, H, W)] = dist;
			}
}

/* adcensus.cu:280-322.  out is (4,H,W): exclusive arm end-points
 * dir 0 left (xx), 1 right (xx), 2 up (yy), 3 down (yy). */
void orc_cross(const float *img, float *out, int H, int W, int L1, float tau1)
{
#pragma omp parallel for collapse(2)
	for (int dir = 0; dir < 4; dir++)
		for (int y = 0; y < H; y++)
			for (int x = 0; x < W; x++) {
				int dx = 0, dy = 0;
				if (dir == 0) dx = -1;
				else if (dir == 1) dx = 1;
				else if (dir == 2) dy = -1;
				else dy = 1;
				int xx, yy;
				int ind1 = y * W + x;
				for (xx = x + dx, yy = y + dy;; xx += dx, yy += dy) {
					if (xx < 0 || xx >= W || yy < 0 || yy >= H) break;     /* :307 */
					int dist = abs(xx - x) > abs(yy - y) ? abs(xx - x) : abs(yy - y);
					if (dist == 1) continue;                                /* :310 */
					int ind2 = yy * W + xx;
					if (fabsf(img[ind1] - img[ind2]) >= tau1) break;        /* :315 rule 1 */
					if (dist >= L1) break;                                  /* :318 rule 2 */
				}
				out[IDX3(dir, y, x, H, W)] = dir <= 1 ? xx : yy;        /* :320 */
			}
}

/* adcensus.cu:343-377.  x0c is always the left image's arms, x1c the right's
 * (main.lua:999,1036). */
void orc_cbca(const float *x0c, const float *x1c, const float *vol, float *out, int D, int H, int W, int direction)
{
#pragma omp parallel for collapse(2)
	for (int d = 0; d < D; d++)
		for (int y = 0; y < H; y++)
			for (int x = 0; x < W; x++) {
				long id = IDX3(d, y, x, H, W);
				int xs = x + d * direction;
				if (xs < 0 || xs >= W) {
					out[id] = vol[id];                                      /* :353-354 */
					continue;
				}
				float sum = 0;
				int cnt = 0;
				int yy_s = (int)fmaxf(x0c[IDX3(2, y, x, H, W)], x1c[IDX3(2, y, xs, H, W)]); /* :359 */
				int yy_t = (int)fminf(x0c[IDX3(3, y, x, H, W)], x1c[IDX3(3, y, xs, H, W)]); /* :360 */
				for (int yy = yy_s + 1; yy < yy_t; yy++) {
					int xx_s = (int)fmaxf(x0c[IDX3(0, yy, x, H, W)], x1c[IDX3(0, yy, xs, H, W)] - d * direction); /* :362 */
					int xx_t = (int)fminf(x0c[IDX3(1, yy, x, H, W)], x1c[IDX3(1, yy, xs, H, W)] - d * direction); /* :363 */
					for (int xx = xx_s + 1; xx < xx_t; xx++) {
						sum += vol[IDX3(d, yy, xx, H, W)];                  /* :367 */
						cnt++;
					}
				}
				out[id] = sum / cnt;                                        /* :373 */
			}
}

/*
 * adcensus.cu:535-618 (kernel sgm2<dir>, one launch per scan step) and 620-697
 * (host: 4 loops of launches in the order right, left, down, up).  Layout
 * (H,W,D): INDEX(0,y,x,d) = (y*W + x)*D + d (adcensus.cu:531-533 with
 * size1=H, size2=W, size3=D).  tmp is indexed d*W + line (:570,576,617), so
 * the horizontal passes need H <= W.  out is accumulated into (+=).
 */
/* image_rows > 0: the volume is a ROW band of an image with image_rows rows and the vertical scans continue across
 * bands -- `tmp` carries the line state in and out, and "first pixel of the line" is decided in image coordinates */
static void sgm2_step_ex(int sgm_direction, int line, int step,
		      const float *x0, const float *x1, const float *in, float *out, float *tmp,
		      int H, int W, int D, float pi1, float pi2, float tau_so, float alpha1,
		      float sgm_q1, float sgm_q2, int direction, int Wt, int yoff, int xoff, int image_rows);

static void sgm2_step(int sgm_direction, int line, int step,
		      const float *x0, const float *x1, const float *in, float *out, float *tmp,
		      int H, int W, int D, float pi1, float pi2, float tau_so, float alpha1,
		      float sgm_q1, float sgm_q2, int direction, int Wt, int yoff, int xoff)
{
	sgm2_step_ex(sgm_direction, line, step, x0, x1, in, out, tmp, H, W, D, pi1, pi2, tau_so, alpha1, sgm_q1, sgm_q2, direction,
		     Wt, yoff, xoff, 0);
}

static void sgm2_step_ex(int sgm_direction, int line, int step,
		      const float *x0, const float *x1, const float *in, float *out, float *tmp,
		      int H, int W, int D, float pi1, float pi2, float tau_so, float alpha1,
		      float sgm_q1, float sgm_q2, int direction, int Wt, int yoff, int xoff, int image_rows)
{
	/* (H,W,D) may be a band of the Ht x Wt image: volume pixel (y,x) is image pixel (yoff+y, xoff+x);
	 * a band holds whole scanlines of the passes run on it, so the first-pixel test stays in volume
	 * coordinates while image look-ups use image coordinates */
	int x, y, dx, dy;
	if (sgm_direction == 0) { x = step; y = line; dx = 1; dy = 0; }
	else if (sgm_direction == 1) { x = W - 1 - step; y = line; dx = -1; dy = 0; }
	else if (sgm_direction == 2) { x = line; y = step; dx = 0; dy = 1; }
	else { x = line; y = H - 1 - step; dx = 0; dy = -1; }

	long base = ((long)y * W + x) * D;
	int first = (y - dy < 0 || y - dy >= H || x - dx < 0 || x - dx >= W);   /* :567-572 */
	if (image_rows > 0 && dy != 0) first = (y + yoff - dy < 0 || y + yoff - dy >= image_rows);
	if (first) {
		for (int d = 0; d < D; d++) {
			float val = in[base + d];
			out[base + d] += val;
			tmp[(long)d * W + line] = val;
		}
		return;
	}

	float output_s[512], output_min[512];
	for (int d = 0; d < D; d++) output_s[d] = output_min[d] = tmp[(long)d * W + line]; /* :576 */
	for (int i = 256; i > 0; i /= 2)                                     /* :579-584 */
		for (int d = 0; d < i; d++)
			if (d + i < D && output_min[d + i] < output_min[d]) output_min[d] = output_min[d + i];

	int ind2 = (y + yoff) * Wt + (x + xoff);
	float D1 = fabsf(x0[ind2] - x0[ind2 - dy * Wt - dx]);                /* :587 */
	for (int d = 0; d < D; d++) {
		float D2;
		int xx = x + xoff + d * direction;
		if (xx < 0 || xx >= Wt || xx - dx < 0 || xx - dx >= Wt) D2 = 10; /* :590-591 */
		else D2 = fabsf(x1[ind2 + d * direction] - x1[ind2 + d * direction - dy * Wt - dx]); /* :593 */
		float P1, P2;
		if (D1 < tau_so && D2 < tau_so) { P1 = pi1; P2 = pi2; }          /* :596-598 */
		else if (D1 > tau_so && D2 > tau_so) { P1 = pi1 / (sgm_q1 * sgm_q2); P2 = pi2 / (sgm_q1 * sgm_q2); }
		else { P1 = pi1 / sgm_q1; P2 = pi2 / sgm_q1; }

		float cost = fminf(output_s[d], output_min[0] + P2);              /* :607 */
		if (d - 1 >= 0) cost = fminf(cost, output_s[d - 1] + (sgm_direction == 2 ? P1 / alpha1 : P1)); /* :609 */
		if (d + 1 < D) cost = fminf(cost, output_s[d + 1] + (sgm_direction == 3 ? P1 / alpha1 : P1));  /* :612 */

		float val = in[base + d] + cost - output_min[0];                  /* :615 */
		out[base + d] += val;                                             /* :616 */
		tmp[(long)d * W + line] = val;                                    /* :617 */
	}
}

/* the passes selected by pass_mask (bit sd) over a band volume; tmp holds W*D floats, indexed d*W + line
 * like the reference (adcensus.cu:570): horizontal passes therefore need H <= W (row bands), vertical
 * passes have line < W by construction */
void orc_sgm2_band(const float *x0, const float *x1, const float *in, float *out, float *tmp,
		   int H, int W, int D, int Wt, int yoff, int xoff, float pi1, float pi2, float tau_so, float alpha1,
		   float sgm_q1, float sgm_q2, int direction, int pass_mask)
{
	for (int sd = 0; sd < 4; sd++) {
		if (!(pass_mask & (1 << sd))) continue;
		int nlines = sd < 2 ? H : W;
		int nsteps = sd < 2 ? W : H;
		/* lines are independent; steps are sequential (one launch each, :639-693) */
#pragma omp parallel for
		for (int line = 0; line < nlines; line++)
			for (int step = 0; step < nsteps; step++)
				sgm2_step(sd, line, step, x0, x1, in, out, tmp, H, W, D, pi1, pi2, tau_so,
					  alpha1, sgm_q1, sgm_q2, direction, Wt, yoff, xoff);
	}
}

/* one vertical pass (sd 2 down / 3 up) over the rows [yoff, yoff + H) of an image with Ht rows, columns [xa, xb); the line
 * state enters and leaves through tmp (D x W, d * W + line): the CPU statement of mccnn_sgm2_rows' chained vertical scans */
void orc_sgm2_vrows(const float *x0, const float *x1, const float *in, float *out, float *tmp,
		    int H, int W, int D, int Ht, int yoff, float pi1, float pi2, float tau_so, float alpha1,
		    float sgm_q1, float sgm_q2, int direction, int sd, int xa, int xb)
{
#pragma omp parallel for
	for (int line = xa; line < xb; line++)
		for (int step = 0; step < H; step++)
			sgm2_step_ex(sd, line, step, x0, x1, in, out, tmp, H, W, D, pi1, pi2, tau_so, alpha1, sgm_q1, sgm_q2, direction,
				     W, yoff, 0, Ht);
}

void orc_sgm2(const float *x0, const float *x1, const float *in, float *out, float *tmp,
	      int H, int W, int D, float pi1, float pi2, float tau_so, float alpha1,
	      float sgm_q1, float sgm_q2, int direction)
{
	orc_sgm2_band(x0, x1, in, out, tmp, H, W, D, W, 0, 0, pi1, pi2, tau_so, alpha1, sgm_q1, sgm_q2, direction, 15);
}

/* adcensus.cu:244-262: 1-based argmin over dim 1, strict <, init +inf (NaN skipped) */
void orc_spatial_argmin(const float *in, float *out, int D, int HW)
{
#pragma omp parallel for
	for (int p = 0; p < HW; p++) {
		int argmin = 0;
		float mn = INFINITY;
		for (int i = 0; i < D; i++) {
			float val = in[(long)i * HW + p];
			if (val < mn) { mn = val; argmin = i; }
		}
		out[p] = argmin + 1;
	}
}

/* adcensus.cu:878-899 */
void orc_outlier_detection(const float *d0, const float *d1, float *outlier, int H, int W, int disp_max)
{
#pragma omp parallel for
	for (int id = 0; id < H * W; id++) {
		int x = id % W;
		int d0i = (int)d0[id];
		if (x - d0i < 0) {
			outlier[id] = 1;
		} else if ((double)fabsf(d0[id] - d1[id - d0i]) < 1.1) {
			outlier[id] = 0;
		} else {
			outlier[id] = 1;
			for (int d = 0; d < disp_max; d++)
				if (x - d >= 0 && (double)fabsf((float)d - d1[id - d]) < 1.1) {
					outlier[id] = 2;
					break;
				}
		}
	}
}

/* adcensus.cu:1079-1105 */
void orc_interpolate_occlusion(const float *d0, const float *outlier, float *out, int H, int W)
{
#pragma omp parallel for
	for (int id = 0; id < H * W; id++) {
		if (outlier[id] != 1) { out[id] = d0[id]; continue; }
		int x = id % W;
		int dx = 0;
		while (x + dx >= 0 && outlier[id + dx] != 0) dx--;
		if (x + dx < 0) {
			dx = 0;
			while (x + dx < W && outlier[id + dx] != 0) dx++;
		}
		out[id] = (x + dx < W) ? d0[id + dx] : d0[id];
	}
}

/* adcensus.cu:1001-1058 */
void orc_interpolate_mismatch(const float *d0, const float *outlier, float *out, int H, int W)
{
	static const float dir[] = {
		0, 1, -0.5, 1, -1, 1, -1, 0.5, -1, 0, -1, -0.5, -1, -1, -0.5, -1,
		0, -1, 0.5, -1, 1, -1, 1, -0.5, 1, 0, 1, 0.5, 1, 1, 0.5, 1};
#pragma omp parallel for
	for (int id = 0; id < H * W; id++) {
		if (outlier[id] != 2) { out[id] = d0[id]; continue; }
		float vals[16];
		int n = 0;
		int x = id % W, y = id / W;
		for (int d = 0; d < 16; d++) {
			float dx = dir[2 * d], dy = dir[2 * d + 1];
			float xx = x, yy = y;
			int xx_i = (int)roundf(xx), yy_i = (int)roundf(yy);
			while (0 <= yy_i && yy_i < H && 0 <= xx_i && xx_i < W && outlier[yy_i * W + xx_i] == 2) {
				xx += dx;
				yy += dy;
				xx_i = (int)roundf(xx);
				yy_i = (int)roundf(yy);
			}
			if (0 <= yy_i && yy_i < H && 0 <= xx_i && xx_i < W) vals[n++] = d0[yy_i * W + xx_i];
		}
		sel_sort(vals, n);
		out[id] = vals[n / 2];                                              /* :1056 */
	}
}

/* adcensus.cu:1205-1220 */
void orc_subpixel_enchancement(const float *d0, const float *vol, float *out, int H, int W, int disp_max)
{
	long HW = (long)H * W;
#pragma omp parallel for
	for (long id = 0; id < HW; id++) {
		int d = (int)d0[id];
		out[id] = d;
		if (1 <= d && d < disp_max - 1) {
			float cn = vol[(d - 1) * HW + id];
			float cz = vol[d * HW + id];
			float cp = vol[(d + 1) * HW + id];
			float denom = 2 * (cp + cn - 2 * cz);
			if ((double)denom > 1e-5) {
				double q = (double)((cp - cn) / denom);
				out[id] = (float)(d - fmin(1.0, fmax(-1.0, q)));
			}
		}
	}
}

/* adcensus.cu:1575-1594 */
void orc_median2d(const float *img, float *out, int H, int W, int ksize)
{
	int r = ksize / 2;
#pragma omp parallel for
	for (int id = 0; id < H * W; id++) {
		int x = id % W, y = id / W;
		float xs[11 * 11];
		int n = 0;
		for (int xx = x - r; xx <= x + r; xx++)
			for (int yy = y - r; yy <= y + r; yy++)
				if (0 <= xx && xx < W && 0 <= yy && yy < H) xs[n++] = img[yy * W + xx];
		sel_sort(xs, n);
		out[id] = xs[n / 2];
	}
}

/* adcensus.cu:1241-1261; r = ksize/2 (host :1275).  The weight index i
 * advances on every tap, in-bounds or not (:1252). */
void orc_mean2d(const float *img, const float *kernel, float *out, int H, int W, int ksize, float alpha2)
{
	int r = ksize / 2;
#pragma omp parallel for
	for (int id = 0; id < H * W; id++) {
		int x = id % W, y = id / W;
		float sum = 0, cnt = 0;
		int i = 0;
		for (int xx = x - r; xx <= x + r; xx++)
			for (int yy = y - r; yy <= y + r; yy++, i++)
				if (0 <= xx && xx < W && 0 <= yy && yy < H && fabsf(img[yy * W + xx] - img[y * W + x]) < alpha2) {
					sum = fmaf(img[yy * W + xx], kernel[i], sum);          /* :1254 */
					cnt += kernel[i];                                      /* :1255 */
				}
		out[id] = sum / cnt;
	}
}

/* main.lua:528-540: un-normalised Gaussian, computed in double, stored as float */
int orc_gaussian(double sigma, float *out)
{
	int kr = (int)ceil(sigma * 3);
	int ks = kr * 2 + 1;
	if (out)
		for (int i = 0; i < ks; i++)
			for (int j = 0; j < ks; j++) {
				int y = i - kr, x = j - kr;
				out[i * ks + j] = (float)exp(-(double)(x * x + y * y) / (2 * sigma * sigma));
			}
	return ks;
}

/* main.lua:1008: (1,D,H,W) -> contiguous (1,H,W,D) */
void orc_transpose_dhw_to_hwd(const float *in, float *out, int D, int H, int W)
{
	long HW = (long)H * W;
#pragma omp parallel for
	for (long p = 0; p < HW; p++)
		for (int d = 0; d < D; d++) out[p * D + d] = in[d * HW + p];
}

/* main.lua:1020: vol(1,D,H,W) = out(1,H,W,D) permuted, / 4 */
void orc_transpose_hwd_to_dhw_div4(const float *in, float *out, int D, int H, int W)
{
	long HW = (long)H * W;
#pragma omp parallel for
	for (long p = 0; p < HW; p++)
		for (int d = 0; d < D; d++) out[d * HW + p] = in[p * D + d] / 4;
}

/* main.lua:929-1082, arch == 'fast', from the tower output onwards */
int orc_stereo_predict(const float *featL, const float *featR, const float *imgL, const float *imgR,
		       int C, int D, int H, int W, const orc_params *p,
		       float *disp_out, float *volL_out, float *volR_out)
{
	if (C < 1 || D < 1 || D > 512 || H < 1 || W < 1 || H > W) return -1;
	long HW = (long)H * W;
	long V = (long)D * HW;
	float *vols = (float *)malloc(2 * V * sizeof(float));
	float *tmpv = (float *)malloc(V * sizeof(float));
	float *outv = (float *)malloc(V * sizeof(float));
	float *x0c = (float *)malloc(4 * HW * sizeof(float));
	float *x1c = (float *)malloc(4 * HW * sizeof(float));
	float *sgm_tmp = (float *)malloc((long)W * D * sizeof(float));
	float *disp[2];
	disp[0] = (float *)malloc(HW * sizeof(float));
	disp[1] = (float *)malloc(HW * sizeof(float));

	for (long i = 0; i < 2 * V; i++) vols[i] = NAN;                      /* :946 */
	orc_stereo_join(featL, featR, vols, vols + V, C, D, H, W);          /* :947 */
	orc_fix_border(vols, D, H, W, p->border, -1);                        /* :948 */
	orc_fix_border(vols + V, D, H, W, p->border, 1);                     /* :949 */

	const int directions[2] = {1, -1};                                   /* :955 */
	for (int k = 0; k < 2; k++) {
		int direction = directions[k];
		float *vol = direction == -1 ? vols : vols + V;                  /* :986 */
		orc_cross(imgL, x0c, H, W, p->L1, p->tau1);                      /* :995 */
		orc_cross(imgR, x1c, H, W, p->L1, p->tau1);                      /* :996 */
		for (int i = 0; i < p->cbca_i1; i++) {                           /* :998-1001 */
			orc_cbca(x0c, x1c, vol, tmpv, D, H, W, direction);
			memcpy(vol, tmpv, V * sizeof(float));
		}
		/* :1008-1020 */
		orc_transpose_dhw_to_hwd(vol, tmpv, D, H, W);
		for (int it = 0; it < p->sgm_i; it++) {
			memset(outv, 0, V * sizeof(float));
			orc_sgm2(imgL, imgR, tmpv, outv, sgm_tmp, H, W, D, p->pi1, p->pi2, p->tau_so,
				 p->alpha1, p->sgm_q1, p->sgm_q2, direction);
			for (long i = 0; i < V; i++) tmpv[i] = outv[i] / 4;          /* :1017 */
		}
		orc_transpose_hwd_to_dhw_div4(outv, vol, D, H, W);              /* :1019-1020 */
		for (int i = 0; i < p->cbca_i2; i++) {                           /* :1035-1038 */
			orc_cbca(x0c, x1c, vol, tmpv, D, H, W, direction);
			memcpy(vol, tmpv, V * sizeof(float));
		}
		float *dd = disp[direction == 1 ? 0 : 1];                        /* :1049-1050 */
		orc_spatial_argmin(vol, dd, D, (int)HW);
		for (long i = 0; i < HW; i++) dd[i] -= 1;
	}
	if (volL_out) memcpy(volL_out, vols, V * sizeof(float));
	if (volR_out) memcpy(volR_out, vols + V, V * sizeof(float));

	/* one buffer per stage keeps the data flow of main.lua:1054-1079 explicit */
	float *stage[6];
	for (int i = 0; i < 6; i++) stage[i] = (float *)malloc(HW * sizeof(float));
	const float *cur = disp[1];                                          /* disp[2] in Lua: left map */
	if (p->lr_check) {                                                   /* :1054-1066 */
		float *outlier = stage[0];
		orc_outlier_detection(disp[1], disp[0], outlier, H, W, D);       /* :1056 */
		orc_interpolate_occlusion(cur, outlier, stage[1], H, W);         /* :1058 */
		orc_interpolate_mismatch(stage[1], outlier, stage[2], H, W);     /* :1063 */
		cur = stage[2];
	}
	orc_subpixel_enchancement(cur, vols /* left volume */, stage[3], H, W, D); /* :1068 */
	orc_median2d(stage[3], stage[4], H, W, 5);                           /* :1073 */
	int ks = orc_gaussian(p->blur_sigma, NULL);
	float *kern = (float *)malloc((long)ks * ks * sizeof(float));
	orc_gaussian(p->blur_sigma, kern);
	orc_mean2d(stage[4], kern, stage[5], H, W, ks, p->blur_t);           /* :1078 */
	memcpy(disp_out, stage[5], HW * sizeof(float));
	for (int i = 0; i < 6; i++) free(stage[i]);

	free(kern); free(vols); free(tmpv); free(outv); free(x0c); free(x1c); free(sgm_tmp);
	free(disp[0]); free(disp[1]);
	return 0;
}

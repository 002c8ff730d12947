/*
 * adcensus_oracle.h -- TEST INFRASTRUCTURE, not product code.
 *
 * CPU restatement (plain C, fp32) of the reference's stereo-method operators
 * (jzbontar/mc-cnn adcensus.cu) and of the Lua orchestration that chains them
 * (main.lua:922-1082).  It is the checker for the CUDA path: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline leg may call it.
 *
 * Parity status: PINNED.  tests/golden/ holds outputs of the reference's own
 * kernels (oracle/_ref/libadcensus_ref.so = adcensus.cu compiled unmodified)
 * run on a B200 over seeded inputs; tests/test_oracle_golden.py checks this
 * restatement against them bit for bit.
 *
 * All tensors are contiguous float32, shapes as in the reference:
 *   features (C,H,W)   volumes (D,H,W)   SGM volumes (H,W,D)   images (H,W)
 */
#ifndef ADCENSUS_ORACLE_H
#define ADCENSUS_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

/* stereo-method hyper-parameters (main.lua:70-295) */
typedef struct orc_params {
	int L1;          /* cross arm length            */
	float tau1;      /* cross colour threshold      */
	int cbca_i1;     /* CBCA iterations before SGM  */
	int cbca_i2;     /* CBCA iterations after SGM   */
	float pi1, pi2;  /* SGM penalties               */
	float sgm_q1, sgm_q2;
	float alpha1;
	float tau_so;
	int sgm_i;       /* SGM iterations (1 in every preset) */
	double blur_sigma;
	float blur_t;
	int border;      /* fix_border n = (window-1)/2 (main.lua:923); 4 for 4x conv3 */
	int lr_check;    /* 1 for kitti/kitti2015 (main.lua:1054), 0 for mb */
} orc_params;

void orc_normalize_forward(const float *in, float *norm, float *out, int N, int C, int H, int W);
void orc_stereo_join(const float *L, const float *R, float *outL, float *outR, int C, int D, int H, int W);
void orc_fix_border(float *vol, int D, int H, int W, int n, int direction);
void orc_ad(const float *x0, const float *x1, float *out, int D, int H, int W, int direction);
void orc_census(const float *x0, const float *x1, float *out, int D, int nch, int H, int W, int direction);
void orc_cross(const float *img, float *out, int H, int W, int L1, float tau1);
void orc_cbca(const float *x0c, const float *x1c, const float *vol, float *out, int D, int H, int W, int direction);
void orc_sgm2(const float *x0, const float *x1, const float *in, float *out, float *tmp,
	      int H, int W, int D, float pi1, float pi2, float tau_so, float alpha1,
	      float sgm_q1, float sgm_q2, int direction);
void orc_sgm2_band(const float *x0, const float *x1, const float *in, float *out, float *tmp,
		   int H, int W, int D, int Wt, int yoff, int xoff, float pi1, float pi2, float tau_so, float alpha1,
		   float sgm_q1, float sgm_q2, int direction, int pass_mask);
void orc_sgm2_vrows(const float *x0, const float *x1, const float *in, float *out, float *tmp,
		    int H, int W, int D, int Ht, int yoff, float pi1, float pi2, float tau_so, float alpha1,
		    float sgm_q1, float sgm_q2, int direction, int sd, int xa, int xb);
void orc_spatial_argmin(const float *in, float *out, int D, int HW);
void orc_outlier_detection(const float *d0, const float *d1, float *outlier, int H, int W, int disp_max);
void orc_interpolate_occlusion(const float *d0, const float *outlier, float *out, int H, int W);
void orc_interpolate_mismatch(const float *d0, const float *outlier, float *out, int H, int W);
void orc_subpixel_enchancement(const float *d0, const float *vol, float *out, int H, int W, int disp_max);
void orc_median2d(const float *img, float *out, int H, int W, int ksize);
void orc_mean2d(const float *img, const float *kernel, float *out, int H, int W, int ksize, float alpha2);
/* main.lua:528-540; returns ksize, writes ksize*ksize floats if out != NULL */
int orc_gaussian(double sigma, float *out);
void orc_transpose_dhw_to_hwd(const float *in, float *out, int D, int H, int W);
void orc_transpose_hwd_to_dhw_div4(const float *in, float *out, int D, int H, int W);

/*
 * main.lua:929-1082 for arch == 'fast', starting from the tower output:
 * featL/featR (C,H,W) unit-norm features, imgL/imgR (H,W) standardised images.
 * disp (H,W) is the returned left disparity map.  volL/volR (D,H,W), when not
 * NULL, receive what `-a predict` writes to left.bin/right.bin.
 * Returns 0, or -1 on bad arguments.
 */
int orc_stereo_predict(const float *featL, const float *featR, const float *imgL, const float *imgR,
		       int C, int D, int H, int W, const orc_params *p,
		       float *disp, float *volL, float *volR);

#ifdef __cplusplus
}
#endif
#endif

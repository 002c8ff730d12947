// pipeline.cu -- stereo_predict (main.lua:929-1082, arch 'fast') as one native
// object: every device buffer is allocated once at create time, a run is a fixed
// sequence of kernel launches on one stream (no allocation, no host sync), so a
// batch driver can keep one pipeline per GPU / per stream busy back to back.
//
// Stage order, which volume is "left"/"right", the direction loop {+1,-1}, the /4
// and the permutes follow main.lua line by line (cited inline); the kernels are
// the ones behind the adcensus_* C ABI.
#include <math.h>
#include <new>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

// internal entry points of the other translation units
size_t adc_packed_words(int H, int W);
int adc_pack_arms(const float *xc, uint32_t *pk, int which, int H, int W, int *maxlen_dev, cudaStream_t s);
int adc_cbca_packed(const uint32_t *pk, const float *x0c, const float *x1c,
		    const float *vol, float *out, int D, int H, int W, int direction, int maxlen, cudaStream_t s, int fast);
size_t adc_sgm_table_bytes(int H, int W, int D);
int adc_sgm2(const float *x0, const float *x1, const float *in, float *out, uint8_t *tab, int H, int W, int D,
	     float pi1, float pi2, float tau_so, float alpha1, float q1, float q2, int direction,
	     bool zero_out, cudaStream_t s);

struct mccnn_pipeline {
	int C, D, H, W, device;
	mccnn_params prm;
	long HW, V;
	size_t bytes;
	int launches;
	int fast_cbca;   // opt-in approximate CBCA (mccnn_pipeline_set_fast_cbca); 0 = exact (default)
	// device buffers
	float *vols;      // 2V: [0] left volume, [1] right volume (main.lua:946)
	float *bufA;      // V : CBCA ping-pong / SGM transposed input
	float *bufC;      // V : SGM accumulator (H,W,D)
	float *x0c, *x1c; // 4HW each: cross arms (main.lua:993-996)
	uint32_t *packed; // packed arm lengths of both images (adc_packed_words)
	int *maxlen;
	float *maps;      // 8HW: disparity maps and stage outputs
	float *gauss;     // ks*ks
	uint8_t *sgmtab;  // SGM penalty-class tables
	int ks;
	// staging for the host-buffer entry points: two slots so that the copy of pair i+1 overlaps
	// the kernels of pair i
	float *h_feat[2], *h_img[2], *h_disp[2];  // device copies per slot: 2F, 2HW, HW
	cudaStream_t own_stream, copy_stream, out_stream;
	cudaEvent_t ev_in[2], ev_free[2], ev_out[2];
};

namespace {

int dev_alloc(void **p, size_t bytes, size_t *acc)
{
	cudaError_t e = cudaMalloc(p, bytes);
	if (e != cudaSuccess) return (int)e;
	*acc += bytes;
	return 0;
}

struct DeviceGuard {
	int prev;
	explicit DeviceGuard(int dev) { cudaGetDevice(&prev); cudaSetDevice(dev); }
	~DeviceGuard() { cudaSetDevice(prev); }
};

}  // namespace

extern "C" int mccnn_pipeline_create(mccnn_pipeline **out, int C, int D, int H, int W, const mccnn_params *params, int device)
{
	if (!out || !params || C < 1 || D < 1 || H < 1 || W < 1) return ADCENSUS_EINVAL;
	if (C > 128 || D > ADCENSUS_MAX_DISP) return ADCENSUS_ELIMIT;
	if (params->cbca_i1 < 0 || params->cbca_i2 < 0 || params->sgm_i < 0 || params->border < 0 || params->border + 1 > W)
		return ADCENSUS_EINVAL;
	DeviceGuard g(device);
	mccnn_pipeline *p = new (std::nothrow) mccnn_pipeline();
	if (!p) return ADCENSUS_EINVAL;
	memset(p, 0, sizeof(*p));
	p->C = C; p->D = D; p->H = H; p->W = W; p->device = device;
	p->prm = *params;
	p->HW = (long)H * W;
	p->V = (long)D * p->HW;
	int rc = 0;
	const size_t f = sizeof(float);
	if (!rc) rc = dev_alloc((void **)&p->vols, 2 * p->V * f, &p->bytes);
	if (!rc) rc = dev_alloc((void **)&p->bufA, p->V * f, &p->bytes);
	if (!rc) rc = dev_alloc((void **)&p->bufC, p->V * f, &p->bytes);
	if (!rc) rc = dev_alloc((void **)&p->x0c, 4 * p->HW * f, &p->bytes);
	if (!rc) rc = dev_alloc((void **)&p->x1c, 4 * p->HW * f, &p->bytes);
	if (!rc) rc = dev_alloc((void **)&p->packed, adc_packed_words(H, W) * sizeof(uint32_t), &p->bytes);
	if (!rc) rc = dev_alloc((void **)&p->maxlen, sizeof(int), &p->bytes);
	if (!rc) rc = dev_alloc((void **)&p->maps, 8 * p->HW * f, &p->bytes);
	if (!rc) rc = dev_alloc((void **)&p->sgmtab, adc_sgm_table_bytes(H, W, D), &p->bytes);
	p->ks = mccnn_gaussian(params->blur_sigma, nullptr);
	if (!rc) rc = dev_alloc((void **)&p->gauss, (size_t)p->ks * p->ks * f, &p->bytes);
	if (!rc) {
		float *hk = (float *)malloc((size_t)p->ks * p->ks * f);
		mccnn_gaussian(params->blur_sigma, hk);  // main.lua:1078 gaussian(opt.blur_sigma):cuda()
		rc = (int)cudaMemcpy(p->gauss, hk, (size_t)p->ks * p->ks * f, cudaMemcpyHostToDevice);
		free(hk);
	}
	if (rc) {
		mccnn_pipeline_destroy(p);
		return rc;
	}
	*out = p;
	return 0;
}

extern "C" void mccnn_pipeline_destroy(mccnn_pipeline *p)
{
	if (!p) return;
	DeviceGuard g(p->device);
	cudaFree(p->vols); cudaFree(p->bufA); cudaFree(p->bufC); cudaFree(p->x0c); cudaFree(p->x1c);
	cudaFree(p->packed); cudaFree(p->maxlen); cudaFree(p->maps); cudaFree(p->gauss); cudaFree(p->sgmtab);
	for (int k = 0; k < 2; k++) {
		cudaFree(p->h_feat[k]); cudaFree(p->h_img[k]); cudaFree(p->h_disp[k]);
		if (p->ev_in[k]) cudaEventDestroy(p->ev_in[k]);
		if (p->ev_free[k]) cudaEventDestroy(p->ev_free[k]);
		if (p->ev_out[k]) cudaEventDestroy(p->ev_out[k]);
	}
	if (p->own_stream) cudaStreamDestroy(p->own_stream);
	if (p->copy_stream) cudaStreamDestroy(p->copy_stream);
	if (p->out_stream) cudaStreamDestroy(p->out_stream);
	delete p;
}

extern "C" size_t mccnn_pipeline_device_bytes(const mccnn_pipeline *p) { return p ? p->bytes : 0; }
extern "C" void mccnn_pipeline_set_fast_cbca(mccnn_pipeline *p, int on) { if (p) p->fast_cbca = on ? 1 : 0; }
extern "C" int mccnn_pipeline_launches_per_run(const mccnn_pipeline *p) { return p ? p->launches : 0; }

#define STEP(call)                 \
	do {                       \
		int rc__ = (call); \
		if (rc__) return rc__; \
	} while (0)

extern "C" int mccnn_pipeline_run(mccnn_pipeline *p, const float *featL, const float *featR,
				  const float *imgL, const float *imgR, float *disp,
				  float *volL, float *volR, adcensus_stream_t stream)
{
	if (!p || !featL || !featR || !imgL || !imgR || !disp) return ADCENSUS_EINVAL;
	DeviceGuard g(p->device);
	cudaStream_t s = adc_stream(stream);
	const int C = p->C, D = p->D, H = p->H, W = p->W;
	const long HW = p->HW, V = p->V;
	const mccnn_params &o = p->prm;
	int nl = 0;

	float *volsL = p->vols, *volsR = p->vols + V;
	STEP(mccnn_fill_invalid(volsL, volsR, D, H, W, s)); nl += 1;                                 // main.lua:946 (only what :947 leaves)
	STEP(adcensus_StereoJoin(featL, featR, volsL, volsR, C, D, H, W, s)); nl += 1;               // :947
	STEP(mccnn_fix_border(volsL, D, H, W, o.border, -1, s));                                     // :948
	STEP(mccnn_fix_border(volsR, D, H, W, o.border, 1, s)); nl += o.border ? 2 : 0;              // :949

	// cross arms: identical for both directions (main.lua:993-996 recomputes them)
	const int maxlen = o.L1 > 2 ? o.L1 : 2;  // bound on the arm length cross() can produce
	STEP(adcensus_cross(imgL, p->x0c, H, W, o.L1, o.tau1, s));
	STEP(adcensus_cross(imgR, p->x1c, H, W, o.L1, o.tau1, s));
	STEP(adc_pack_arms(p->x0c, p->packed, 0, H, W, p->maxlen, s));
	STEP(adc_pack_arms(p->x1c, p->packed, 1, H, W, p->maxlen, s)); nl += 4;

	float *dispR = p->maps, *dispL = p->maps + HW;
	float *final_left = nullptr;
	float *spare = p->bufA;
	const int directions[2] = {1, -1};                                                           // :955
	for (int k = 0; k < 2; k++) {
		const int direction = directions[k];
		float *cur = direction == -1 ? volsL : volsR;                                            // :986
		for (int i = 0; i < o.cbca_i1; i++) {                                                    // :998-1001
			STEP(adc_cbca_packed(p->packed, p->x0c, p->x1c, cur, spare, D, H, W, direction, maxlen, s, p->fast_cbca));
			float *t = cur; cur = spare; spare = t; nl += 1;
		}
		for (int it = 0; it < o.sgm_i; it++) {                                                   // :1008-1020
			STEP(mccnn_transpose_dhw_to_hwd(cur, spare, D, H, W, s));                            // :1008
			STEP(adc_sgm2(imgL, imgR, spare, p->bufC, p->sgmtab, H, W, D, o.pi1, o.pi2, o.tau_so, o.alpha1,
				      o.sgm_q1, o.sgm_q2, direction, /*zero_out=*/true, s));                     // :1014-1016
			STEP(mccnn_transpose_hwd_to_dhw_div4(p->bufC, cur, D, H, W, s));                     // :1017-1020
			nl += 7;
		}
		for (int i = 0; i < o.cbca_i2; i++) {                                                    // :1035-1038
			STEP(adc_cbca_packed(p->packed, p->x0c, p->x1c, cur, spare, D, H, W, direction, maxlen, s, p->fast_cbca));
			float *t = cur; cur = spare; spare = t; nl += 1;
		}
		STEP(mccnn_argmin(cur, direction == 1 ? dispR : dispL, D, (int)HW, s)); nl += 1;         // :1049-1050
		float *dst = direction == 1 ? volR : volL;                                               // :1042-1047
		if (dst) {
			STEP((int)cudaMemcpyAsync(dst, cur, V * sizeof(float), cudaMemcpyDeviceToDevice, s));
		}
		if (direction == -1) final_left = cur;
		// the right volume is dead after its argmin: both of its buffers may serve as spares,
		// `spare` already points at a free one
	}

	float *m = p->maps;
	const float *curd = dispL;                                                                   // disp[2]
	if (o.lr_check) {                                                                            // :1054-1066
		float *outlier = m + 2 * HW;
		STEP(adcensus_outlier_detection(dispL, dispR, outlier, H, W, D, s));                     // :1056
		STEP(adcensus_interpolate_occlusion(curd, outlier, m + 3 * HW, H, W, s));                // :1058
		STEP(adcensus_interpolate_mismatch(m + 3 * HW, outlier, m + 4 * HW, H, W, s));           // :1063
		curd = m + 4 * HW; nl += 3;
	}
	STEP(adcensus_subpixel_enchancement(curd, final_left, m + 5 * HW, H, W, D, s));              // :1068
	STEP(adcensus_median2d(m + 5 * HW, m + 6 * HW, H, W, 5, s));                                 // :1073
	STEP(adcensus_mean2d(m + 6 * HW, p->gauss, disp, H, W, p->ks, o.blur_t, s)); nl += 3;        // :1078
	p->launches = nl;
	return 0;
}

static int host_staging(mccnn_pipeline *p)
{
	if (p->own_stream) return 0;
	const size_t F = (size_t)p->C * p->HW * sizeof(float), I = (size_t)p->HW * sizeof(float);
	int rc = 0;
	for (int k = 0; k < 2 && !rc; k++) {
		if (!rc) rc = dev_alloc((void **)&p->h_feat[k], 2 * F, &p->bytes);
		if (!rc) rc = dev_alloc((void **)&p->h_img[k], 2 * I, &p->bytes);
		if (!rc) rc = dev_alloc((void **)&p->h_disp[k], I, &p->bytes);
		if (!rc) rc = (int)cudaEventCreateWithFlags(&p->ev_in[k], cudaEventDisableTiming);
		if (!rc) rc = (int)cudaEventCreateWithFlags(&p->ev_free[k], cudaEventDisableTiming);
		if (!rc) rc = (int)cudaEventCreateWithFlags(&p->ev_out[k], cudaEventDisableTiming);
	}
	if (!rc) rc = (int)cudaStreamCreateWithFlags(&p->copy_stream, cudaStreamNonBlocking);
	if (!rc) rc = (int)cudaStreamCreateWithFlags(&p->out_stream, cudaStreamNonBlocking);
	if (!rc) rc = (int)cudaStreamCreateWithFlags(&p->own_stream, cudaStreamNonBlocking);
	return rc;
}

// n pairs from host memory (pinned for real overlap): pair i's inputs are copied on a copy stream
// into staging slot i%2 while pair i-1 computes; its disparity map is copied back on the copy
// stream while pair i+1 computes (separate H2D / kernel / D2H streams, two staging slots).
// Returns when every result is in its host buffer.
extern "C" int mccnn_pipeline_run_host_batch(mccnn_pipeline *p, int n, const float *const *featL_host,
					     const float *const *featR_host, const float *const *imgL_host,
					     const float *const *imgR_host, float *const *disp_host)
{
	if (!p || n < 0 || (n > 0 && (!featL_host || !featR_host || !imgL_host || !imgR_host || !disp_host))) return ADCENSUS_EINVAL;
	for (int i = 0; i < n; i++)
		if (!featL_host[i] || !featR_host[i] || !imgL_host[i] || !imgR_host[i] || !disp_host[i]) return ADCENSUS_EINVAL;
	DeviceGuard g(p->device);
	STEP(host_staging(p));
	const size_t F = (size_t)p->C * p->HW * sizeof(float), I = (size_t)p->HW * sizeof(float);
	cudaStream_t cs = p->copy_stream, ks = p->own_stream, os = p->out_stream;  // H2D, kernels, D2H
	for (int i = 0; i < n; i++) {
		const int k = i & 1;
		float *fL = p->h_feat[k], *fR = fL + (size_t)p->C * p->HW;
		float *iL = p->h_img[k], *iR = iL + p->HW;
		if (i >= 2) ADC_CUDA(cudaStreamWaitEvent(cs, p->ev_free[k], 0));   // slot's previous pair consumed
		ADC_CUDA(cudaMemcpyAsync(fL, featL_host[i], F, cudaMemcpyHostToDevice, cs));
		ADC_CUDA(cudaMemcpyAsync(fR, featR_host[i], F, cudaMemcpyHostToDevice, cs));
		ADC_CUDA(cudaMemcpyAsync(iL, imgL_host[i], I, cudaMemcpyHostToDevice, cs));
		ADC_CUDA(cudaMemcpyAsync(iR, imgR_host[i], I, cudaMemcpyHostToDevice, cs));
		ADC_CUDA(cudaEventRecord(p->ev_in[k], cs));
		ADC_CUDA(cudaStreamWaitEvent(ks, p->ev_in[k], 0));
		if (i >= 2) ADC_CUDA(cudaStreamWaitEvent(ks, p->ev_out[k], 0));     // slot's previous result copied out
		STEP(mccnn_pipeline_run(p, fL, fR, iL, iR, p->h_disp[k], nullptr, nullptr, ks));
		ADC_CUDA(cudaEventRecord(p->ev_free[k], ks));
		ADC_CUDA(cudaStreamWaitEvent(os, p->ev_free[k], 0));
		ADC_CUDA(cudaMemcpyAsync(disp_host[i], p->h_disp[k], I, cudaMemcpyDeviceToHost, os));
		ADC_CUDA(cudaEventRecord(p->ev_out[k], os));
	}
	ADC_CUDA(cudaStreamSynchronize(os));
	ADC_CUDA(cudaStreamSynchronize(ks));
	ADC_CUDA(cudaStreamSynchronize(cs));
	return 0;
}

extern "C" int mccnn_pipeline_run_host(mccnn_pipeline *p, const float *featL_host, const float *featR_host,
				       const float *imgL_host, const float *imgR_host, float *disp_host)
{
	return mccnn_pipeline_run_host_batch(p, 1, &featL_host, &featR_host, &imgL_host, &imgR_host, &disp_host);
}

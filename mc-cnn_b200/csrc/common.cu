// common.cu -- device properties cache, stream-ordered scratch, version string
#include "common.cuh"

int adc_num_sms()
{
	static int sms[64] = {0};
	int dev = 0;
	if (cudaGetDevice(&dev) != cudaSuccess) return 148;
	if (!sms[dev & 63]) {
		int n = 0;
		if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
		sms[dev & 63] = n;
	}
	return sms[dev & 63];
}

int adc_scratch_alloc(void **p, size_t bytes, cudaStream_t s)
{
	// keep freed scratch in the device's default pool instead of returning it to the OS at every
	// synchronisation (the default release threshold is 0), so repeated op-level calls reuse it
	static bool tuned[64] = {false};
	int dev = 0;
	if (cudaGetDevice(&dev) == cudaSuccess && !tuned[dev & 63]) {
		cudaMemPool_t pool;
		if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
			unsigned long long keep = ~0ULL;
			cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
		}
		tuned[dev & 63] = true;
	}
	return (int)cudaMallocAsync(p, bytes ? bytes : 4, s);
}

int adc_scratch_free(void *p, cudaStream_t s)
{
	return p ? (int)cudaFreeAsync(p, s) : 0;
}

extern "C" const char *adcensus_version(void) { return "libadcensus_b200 0.1.0 (sm_100a)"; }

// lua_face.cu -- the Lua face of libadcensus for sm_100a: `luaopen_libadcensus`.
//
// Drop-in for adcensus.cu:2061-2105: registers the same 31 names into the global
// Lua table `adcensus` (and the two SpatialLogSoftMax entries into `nn`), each a
// lua_CFunction with the reference's positional signature.  The hot-path names
// unmarshal their tensors exactly like the reference (luaT_checkudata "torch.CudaTensor",
// sizes read from the tensors) and call the C ABI of
// libadcensus_b200.so; results that the reference returns as new tensors are
// allocated with THCudaTensor_new + resizeAs (new_tensor_like, adcensus.cu:40-45)
// and pushed with luaT_pushudata.  Non-zero C-ABI return codes become luaL_error,
// like checkCudaError (adcensus.cu:31-36).  Names outside the stereo hot path
// (training, dataset construction, PNG/PFM I/O, dead kernels) are registered so that
// table lookups behave, and raise "not implemented in libadcensus_b200".
//
// Build: against real Torch7 headers (lua.h, luaT.h, THC.h) to get the .so main.lua
// requires, or against oracle/refshim in this repository's tests,
// where the same driver calls adcensus.* in this library and in the reference.
extern "C" {
#include "lua.h"
#include "lualib.h"
#include "lauxlib.h"
}
#include "luaT.h"
#include "THC.h"

#include <cuda_runtime.h>
#include <stdio.h>

#include "adcensus_b200.h"

namespace {

THCState *get_state(lua_State *L)  // adcensus.cu:21-29
{
	lua_getglobal(L, "cutorch");
	lua_getfield(L, -1, "getState");
	lua_call(L, 0, 1);
	THCState *state = (THCState *)lua_touserdata(L, -1);
	lua_pop(L, 2);
	return state;
}

inline THCudaTensor *arg(lua_State *L, int i) { return (THCudaTensor *)luaT_checkudata(L, i, "torch.CudaTensor"); }

inline void check(lua_State *L, int rc, const char *what)
{
	if (rc == 0) return;
	if (rc > 0) luaL_error(L, "%s: %s", what, cudaGetErrorString((cudaError_t)rc));
	luaL_error(L, "%s: %s", what, rc == ADCENSUS_ELIMIT ? "size limit exceeded" : "invalid argument");
}

THCudaTensor *new_like(THCState *s, THCudaTensor *x)
{
	THCudaTensor *y = THCudaTensor_new(s);
	THCudaTensor_resizeAs(s, y, x);
	return y;
}

#define DATA(t) THCudaTensor_data(state, t)
#define SIZE(t, i) ((int)THCudaTensor_size(state, t, i))
// the reference launches on the legacy default stream; cutorch of that era has no per-tensor stream
#define STREAM ((adcensus_stream_t)0)

int l_StereoJoin(lua_State *L)
{
	THCState *state = get_state(L);
	THCudaTensor *iL = arg(L, 1), *iR = arg(L, 2), *oL = arg(L, 3), *oR = arg(L, 4);
	check(L, adcensus_StereoJoin(DATA(iL), DATA(iR), DATA(oL), DATA(oR), SIZE(iL, 1), SIZE(oL, 1), SIZE(oL, 2), SIZE(oL, 3), STREAM), "StereoJoin");
	return 0;
}

int l_cross(lua_State *L)
{
	THCState *state = get_state(L);
	THCudaTensor *x0 = arg(L, 1), *out = arg(L, 2);
	int L1 = (int)luaL_checkinteger(L, 3);
	float tau1 = (float)luaL_checknumber(L, 4);
	check(L, adcensus_cross(DATA(x0), DATA(out), SIZE(out, 2), SIZE(out, 3), L1, tau1, STREAM), "cross");
	return 0;
}

int l_cbca(lua_State *L)
{
	THCState *state = get_state(L);
	THCudaTensor *x0c = arg(L, 1), *x1c = arg(L, 2), *vin = arg(L, 3), *vout = arg(L, 4);
	int direction = (int)luaL_checkinteger(L, 5);
	check(L, adcensus_cbca(DATA(x0c), DATA(x1c), DATA(vin), DATA(vout), SIZE(vout, 1), SIZE(vout, 2), SIZE(vout, 3), direction, STREAM), "cbca");
	return 0;
}

int l_sgm2(lua_State *L)
{
	THCState *state = get_state(L);
	THCudaTensor *x0 = arg(L, 1), *x1 = arg(L, 2), *in = arg(L, 3), *out = arg(L, 4), *tmp = arg(L, 5);
	float pi1 = (float)luaL_checknumber(L, 6), pi2 = (float)luaL_checknumber(L, 7), tau_so = (float)luaL_checknumber(L, 8);
	float alpha1 = (float)luaL_checknumber(L, 9), q1 = (float)luaL_checknumber(L, 10), q2 = (float)luaL_checknumber(L, 11);
	int direction = (int)luaL_checknumber(L, 12);
	check(L, adcensus_sgm2(DATA(x0), DATA(x1), DATA(in), DATA(out), DATA(tmp), SIZE(in, 1), SIZE(in, 2), SIZE(in, 3),
			       pi1, pi2, tau_so, alpha1, q1, q2, direction, STREAM), "sgm2");
	return 0;
}

int l_outlier_detection(lua_State *L)
{
	THCState *state = get_state(L);
	THCudaTensor *d0 = arg(L, 1), *d1 = arg(L, 2), *outlier = arg(L, 3);
	int disp_max = (int)luaL_checkinteger(L, 4);
	int W = SIZE(d0, 3), H = (int)(THCudaTensor_nElement(state, d0) / W);
	check(L, adcensus_outlier_detection(DATA(d0), DATA(d1), DATA(outlier), H, W, disp_max, STREAM), "outlier_detection");
	return 0;
}

int l_interpolate_occlusion(lua_State *L)
{
	THCState *state = get_state(L);
	THCudaTensor *d0 = arg(L, 1), *outlier = arg(L, 2);
	THCudaTensor *out = new_like(state, d0);
	int W = SIZE(out, 3), H = (int)(THCudaTensor_nElement(state, out) / W);
	check(L, adcensus_interpolate_occlusion(DATA(d0), DATA(outlier), DATA(out), H, W, STREAM), "interpolate_occlusion");
	luaT_pushudata(L, out, "torch.CudaTensor");
	return 1;
}

int l_interpolate_mismatch(lua_State *L)
{
	THCState *state = get_state(L);
	THCudaTensor *d0 = arg(L, 1), *outlier = arg(L, 2);
	THCudaTensor *out = new_like(state, d0);
	check(L, adcensus_interpolate_mismatch(DATA(d0), DATA(outlier), DATA(out), SIZE(out, 2), SIZE(out, 3), STREAM), "interpolate_mismatch");
	luaT_pushudata(L, out, "torch.CudaTensor");
	return 1;
}

int l_subpixel_enchancement(lua_State *L)
{
	THCState *state = get_state(L);
	THCudaTensor *d0 = arg(L, 1), *c2 = arg(L, 2);
	int disp_max = (int)luaL_checkinteger(L, 3);
	THCudaTensor *out = new_like(state, d0);
	check(L, adcensus_subpixel_enchancement(DATA(d0), DATA(c2), DATA(out), SIZE(out, 2), SIZE(out, 3), disp_max, STREAM), "subpixel_enchancement");
	luaT_pushudata(L, out, "torch.CudaTensor");
	return 1;
}

int l_median2d(lua_State *L)
{
	THCState *state = get_state(L);
	THCudaTensor *img = arg(L, 1);
	int ks = (int)luaL_checkinteger(L, 2);
	THCudaTensor *out = new_like(state, img);
	check(L, adcensus_median2d(DATA(img), DATA(out), SIZE(out, 2), SIZE(out, 3), ks, STREAM), "median2d");
	luaT_pushudata(L, out, "torch.CudaTensor");
	return 1;
}

int l_mean2d(lua_State *L)
{
	THCState *state = get_state(L);
	THCudaTensor *img = arg(L, 1), *kernel = arg(L, 2);
	float alpha2 = (float)luaL_checknumber(L, 3);
	THCudaTensor *out = new_like(state, img);
	check(L, adcensus_mean2d(DATA(img), DATA(kernel), DATA(out), SIZE(out, 2), SIZE(out, 3), SIZE(kernel, 0), alpha2, STREAM), "mean2d");
	luaT_pushudata(L, out, "torch.CudaTensor");
	return 1;
}

int l_Normalize_forward(lua_State *L)
{
	THCState *state = get_state(L);
	THCudaTensor *in = arg(L, 1), *norm = arg(L, 2), *out = arg(L, 3);
	check(L, adcensus_Normalize_forward(DATA(in), DATA(norm), DATA(out), SIZE(in, 0), SIZE(in, 1), SIZE(in, 2), SIZE(in, 3), STREAM), "Normalize_forward");
	return 0;
}

int l_spatial_argmin(lua_State *L)
{
	THCState *state = get_state(L);
	THCudaTensor *in = arg(L, 1), *out = arg(L, 2);
	check(L, adcensus_spatial_argmin(DATA(in), DATA(out), SIZE(in, 0), SIZE(in, 1), SIZE(in, 2) * SIZE(in, 3), STREAM), "spatial_argmin");
	return 0;
}

int l_ad(lua_State *L)
{
	THCState *state = get_state(L);
	THCudaTensor *x0 = arg(L, 1), *x1 = arg(L, 2), *out = arg(L, 3);
	int direction = (int)luaL_checkinteger(L, 4);
	check(L, adcensus_ad(DATA(x0), DATA(x1), DATA(out), SIZE(out, 1), SIZE(out, 2), SIZE(out, 3), direction, STREAM), "ad");
	return 0;
}

int l_census(lua_State *L)
{
	THCState *state = get_state(L);
	THCudaTensor *x0 = arg(L, 1), *x1 = arg(L, 2), *out = arg(L, 3);
	int direction = (int)luaL_checkinteger(L, 4);
	check(L, adcensus_census(DATA(x0), DATA(x1), DATA(out), SIZE(out, 1), SIZE(x0, 1), SIZE(out, 2), SIZE(out, 3), direction, STREAM), "census");
	return 0;
}

int l_version(lua_State *)
{
	printf("%s\n", adcensus_version());
	return 0;
}

#define NOT_IMPL(name) \
	int l_##name(lua_State *L) { return luaL_error(L, #name ": not implemented in libadcensus_b200 (outside the stereo hot path)"); }
NOT_IMPL(sgm)
NOT_IMPL(sgm3)
NOT_IMPL(copy_fill)
NOT_IMPL(Normalize_backward_input)
NOT_IMPL(Margin2)
NOT_IMPL(StereoL2R)
NOT_IMPL(subset_dataset)
NOT_IMPL(make_dataset)
NOT_IMPL(make_dataset2)
NOT_IMPL(remove_nonvisible)
NOT_IMPL(remove_occluded)
NOT_IMPL(remove_white)
NOT_IMPL(readPNG16)
NOT_IMPL(writePNG16)
NOT_IMPL(writePFM)
NOT_IMPL(grey2jet)
NOT_IMPL(SpatialLogSoftMax_updateOutput)
NOT_IMPL(SpatialLogSoftMax_updateGradInput)

// same order as adcensus.cu:2061-2096
const struct luaL_Reg funcs[] = {
	{"ad", l_ad},
	{"census", l_census},
	{"cross", l_cross},
	{"cbca", l_cbca},
	{"sgm", l_sgm},
	{"sgm2", l_sgm2},
	{"sgm3", l_sgm3},
	{"outlier_detection", l_outlier_detection},
	{"interpolate_occlusion", l_interpolate_occlusion},
	{"interpolate_mismatch", l_interpolate_mismatch},
	{"subpixel_enchancement", l_subpixel_enchancement},
	{"copy_fill", l_copy_fill},
	{"median2d", l_median2d},
	{"mean2d", l_mean2d},
	{"Normalize_forward", l_Normalize_forward},
	{"Normalize_backward_input", l_Normalize_backward_input},
	{"Margin2", l_Margin2},
	{"StereoJoin", l_StereoJoin},
	{"StereoL2R", l_StereoL2R},
	{"subset_dataset", l_subset_dataset},
	{"make_dataset", l_make_dataset},
	{"make_dataset2", l_make_dataset2},
	{"remove_nonvisible", l_remove_nonvisible},
	{"remove_occluded", l_remove_occluded},
	{"remove_white", l_remove_white},
	{"readPNG16", l_readPNG16},
	{"writePNG16", l_writePNG16},
	{"writePFM", l_writePFM},
	{"grey2jet", l_grey2jet},
	{"spatial_argmin", l_spatial_argmin},
	{"version", l_version},
	{NULL, NULL}};

const struct luaL_Reg nn_funcs[] = {
	{"SpatialLogSoftMax_updateOutput", l_SpatialLogSoftMax_updateOutput},
	{"SpatialLogSoftMax_updateGradInput", l_SpatialLogSoftMax_updateGradInput},
	{NULL, NULL}};

}  // namespace

extern "C" int luaopen_libadcensus(lua_State *L)
{
	luaL_openlib(L, "nn", nn_funcs, 0);       // SpatialLogSoftMax.cu:180-189
	luaL_openlib(L, "adcensus", funcs, 0);    // adcensus.cu:2103
	return 1;
}

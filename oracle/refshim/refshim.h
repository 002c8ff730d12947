/*
 * refshim.h -- TEST INFRASTRUCTURE, not product code.
 *
 * A minimal stand-in for the Lua 5.1 C API, luaT and TH/THC (Torch7) headers,
 * just large enough that the reference's operator library
 * (/root/reference/adcensus.cu + SpatialLogSoftMax.cu) compiles UNMODIFIED and
 * that our own Lua face (mc-cnn_b200/csrc/lua_face.cu) compiles in a container
 * that has no LuaJIT / Torch7.  The surface is exactly what those two
 * translation units use (SURVEY.md section 8c lists it).
 *
 * Everything here is declared with C linkage because the reference includes
 * lua.h / lualib.h / lauxlib.h inside an extern "C" block (adcensus.cu:1-5).
 */
#ifndef REFSHIM_H
#define REFSHIM_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- Lua ------------------------------------------------------------- */
typedef struct lua_State lua_State;
typedef int (*lua_CFunction)(lua_State *L);
typedef double lua_Number;
typedef ptrdiff_t lua_Integer;

typedef struct luaL_Reg {
	const char *name;
	lua_CFunction func;
} luaL_Reg;

void lua_getglobal(lua_State *L, const char *name);
void lua_getfield(lua_State *L, int idx, const char *k);
void lua_call(lua_State *L, int nargs, int nresults);
void *lua_touserdata(lua_State *L, int idx);
void lua_pop(lua_State *L, int n);
void lua_pushinteger(lua_State *L, lua_Integer n);

lua_Integer luaL_checkinteger(lua_State *L, int narg);
lua_Number luaL_checknumber(lua_State *L, int narg);
const char *luaL_checkstring(lua_State *L, int narg);
int luaL_error(lua_State *L, const char *fmt, ...);
void luaL_openlib(lua_State *L, const char *libname, const luaL_Reg *l, int nup);

/* ---- luaT ------------------------------------------------------------ */
void *luaT_checkudata(lua_State *L, int ud, const char *tname);
void luaT_pushudata(lua_State *L, void *udata, const char *tname);
int luaT_getfieldcheckboolean(lua_State *L, int ud, const char *field);
double luaT_getfieldchecknumber(lua_State *L, int ud, const char *field);
void *luaT_getfieldcheckudata(lua_State *L, int ud, const char *field, const char *tname);

/* ---- TH / THC -------------------------------------------------------- */
typedef struct THCState { int dummy; } THCState;

#define REFSHIM_MAXDIM 8
#define REFSHIM_TENSOR_FIELDS \
	long size[REFSHIM_MAXDIM]; \
	int nDimension;            \
	void *storage;             \
	int refcount;              \
	int owns;                  \
	int on_device;             \
	int elsize;

typedef struct THCudaTensor { REFSHIM_TENSOR_FIELDS } THCudaTensor;
typedef struct THFloatTensor { REFSHIM_TENSOR_FIELDS } THFloatTensor;
typedef struct THDoubleTensor { REFSHIM_TENSOR_FIELDS } THDoubleTensor;
typedef struct THLongTensor { REFSHIM_TENSOR_FIELDS } THLongTensor;
typedef struct THIntTensor { REFSHIM_TENSOR_FIELDS } THIntTensor;

THCudaTensor *THCudaTensor_new(THCState *state);
void THCudaTensor_resizeAs(THCState *state, THCudaTensor *self, THCudaTensor *src);
float *THCudaTensor_data(THCState *state, const THCudaTensor *self);
long THCudaTensor_size(THCState *state, const THCudaTensor *self, int dim);
long THCudaTensor_nElement(THCState *state, const THCudaTensor *self);
THCudaTensor *THCudaTensor_newContiguous(THCState *state, THCudaTensor *self);
void THCudaTensor_free(THCState *state, THCudaTensor *self);

THFloatTensor *THFloatTensor_new(void);
void THFloatTensor_resizeAs(THFloatTensor *self, THFloatTensor *src);
float *THFloatTensor_data(const THFloatTensor *self);
long THFloatTensor_size(const THFloatTensor *self, int dim);
long THFloatTensor_nElement(const THFloatTensor *self);

double *THDoubleTensor_data(const THDoubleTensor *self);
long THDoubleTensor_size(const THDoubleTensor *self, int dim);
long THDoubleTensor_nElement(const THDoubleTensor *self);

long *THLongTensor_data(const THLongTensor *self);
long THLongTensor_nElement(const THLongTensor *self);

THIntTensor *THIntTensor_new(void);
THIntTensor *THIntTensor_newWithSize1d(long size0);
void THIntTensor_resizeAs(THIntTensor *self, THIntTensor *src);
void THIntTensor_zero(THIntTensor *self);
int *THIntTensor_data(const THIntTensor *self);

void THError(const char *fmt, ...);
void refshim_argcheck(int cond, int argn, const char *msg);
#define THArgCheck(cond, argn, msg) refshim_argcheck((cond), (argn), (msg))

/* ---- driver side (what tests / bench use to call a registered op) ----- */
lua_State *shim_state_new(void);
void shim_state_free(lua_State *L);
/* drop all stack values; tensors the callee allocated are returned to the pool */
void shim_reset(lua_State *L);
void shim_push_cuda_tensor(lua_State *L, void *devptr, int nd, const long *sizes);
void shim_push_float_tensor(lua_State *L, void *hostptr, int nd, const long *sizes);
void shim_push_number(lua_State *L, double v);
void shim_push_string(lua_State *L, const char *s);
/* call lib.fn with the pushed arguments; returns #results (>=0) or -1 on error */
int shim_call(lua_State *L, const char *lib, const char *fn);
const char *shim_last_error(lua_State *L);
/* i-th result (0-based) of the last call */
int shim_result_is_tensor(lua_State *L, int i);
void *shim_result_tensor(lua_State *L, int i, int *nd, long *sizes);
int shim_copy_result_to(lua_State *L, int i, void *dst, size_t bytes);
double shim_result_number(lua_State *L, int i);
int shim_has_function(const char *lib, const char *fn);
int shim_num_functions(const char *lib);
const char *shim_function_name(const char *lib, int i);
/* release cached device blocks */
void shim_pool_trim(void);

#ifdef __cplusplus
}
#endif
#endif

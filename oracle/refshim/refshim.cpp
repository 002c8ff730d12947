/*
 * refshim.cpp -- TEST INFRASTRUCTURE, not product code.
 *
 * Implements the fake Lua stack / luaT / TH(C) surface declared in refshim.h
 * plus the small driver API (shim_*) that tests and bench.py use to invoke a
 * registered lua_CFunction (e.g. the reference's `adcensus.StereoJoin`) on raw
 * device pointers.  Linked into oracle/_ref/libadcensus_ref.so together with
 * the unmodified reference sources, and into the shim build of our own Lua
 * face, so that the same driver exercises either library.
 *
 * luaL_error / THError do not return in real Lua (longjmp); here they throw a
 * C++ exception that shim_call() catches and turns into return code -1.
 */
#include "refshim.h"

#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <map>
#include <stdexcept>
#include <string>
#include <vector>

namespace {

enum Kind { K_NIL, K_NUM, K_STR, K_UDATA, K_DUMMY };

struct Value {
	Kind kind = K_NIL;
	double num = 0;
	std::string str;
	void *ud = nullptr;
	std::string tname;
};

struct ShimError : public std::runtime_error {
	explicit ShimError(const std::string &s) : std::runtime_error(s) {}
};

/* generic view of every TH*Tensor (they share REFSHIM_TENSOR_FIELDS) */
typedef THCudaTensor AnyTensor;

std::map<std::string, std::vector<luaL_Reg>> &registry()
{
	static std::map<std::string, std::vector<luaL_Reg>> r;
	return r;
}

/* size-keyed cache of device blocks, so that ops returning a new tensor
 * (new_tensor_like, adcensus.cu:40-45) do not pay cudaMalloc/cudaFree (and the
 * implicit device sync of cudaFree) on every call -- cutorch's caching
 * allocator plays the same role under real Torch7. */
std::multimap<size_t, void *> &pool()
{
	static std::multimap<size_t, void *> p;
	return p;
}

void *pool_alloc(size_t bytes)
{
	auto it = pool().find(bytes);
	if (it != pool().end()) {
		void *p = it->second;
		pool().erase(it);
		return p;
	}
	void *p = nullptr;
	cudaError_t e = cudaMalloc(&p, bytes ? bytes : 4);
	if (e != cudaSuccess) throw ShimError(std::string("cudaMalloc: ") + cudaGetErrorString(e));
	return p;
}

void pool_free(size_t bytes, void *p) { pool().insert({bytes, p}); }

long numel(const AnyTensor *t)
{
	if (t->nDimension == 0) return 0;
	long n = 1;
	for (int i = 0; i < t->nDimension; i++) n *= t->size[i];
	return n;
}

AnyTensor *tensor_new(int on_device, int elsize)
{
	AnyTensor *t = (AnyTensor *)calloc(1, sizeof(AnyTensor));
	t->refcount = 1;
	t->on_device = on_device;
	t->elsize = elsize;
	return t;
}

void tensor_release_storage(AnyTensor *t)
{
	if (t->owns && t->storage) {
		if (t->on_device) pool_free((size_t)numel(t) * t->elsize, t->storage);
		else free(t->storage);
	}
	t->storage = nullptr;
	t->owns = 0;
}

void tensor_free(AnyTensor *t)
{
	if (!t) return;
	if (--t->refcount > 0) return;
	tensor_release_storage(t);
	free(t);
}

void tensor_resize_as(AnyTensor *self, const AnyTensor *src)
{
	tensor_release_storage(self);
	self->nDimension = src->nDimension;
	memcpy(self->size, src->size, sizeof(self->size));
	size_t bytes = (size_t)numel(self) * self->elsize;
	self->storage = self->on_device ? pool_alloc(bytes) : malloc(bytes ? bytes : 4);
	self->owns = 1;
}

THCState g_state;

}  // namespace

struct lua_State {
	std::vector<Value> stack;
	std::vector<Value> results;
	std::vector<AnyTensor *> wrappers; /* tensors created by shim_push_* */
	std::string error;
};

namespace {

Value &at(lua_State *L, int idx)
{
	int n = (int)L->stack.size();
	int pos = idx > 0 ? idx - 1 : n + idx;
	if (pos < 0 || pos >= n) throw ShimError("bad stack index " + std::to_string(idx));
	return L->stack[pos];
}

std::string vformat(const char *fmt, va_list ap)
{
	char buf[1024];
	vsnprintf(buf, sizeof(buf), fmt, ap);
	return buf;
}

}  // namespace

extern "C" {

/* ---- Lua ------------------------------------------------------------- */
void lua_getglobal(lua_State *L, const char *) { Value v; v.kind = K_DUMMY; L->stack.push_back(v); }
void lua_getfield(lua_State *L, int, const char *) { Value v; v.kind = K_DUMMY; L->stack.push_back(v); }

void lua_call(lua_State *L, int nargs, int nresults)
{
	for (int i = 0; i < nargs + 1 && !L->stack.empty(); i++) L->stack.pop_back();
	for (int i = 0; i < nresults; i++) { Value v; v.kind = K_DUMMY; L->stack.push_back(v); }
}

void *lua_touserdata(lua_State *L, int idx)
{
	Value &v = at(L, idx);
	if (v.kind == K_DUMMY) return &g_state; /* cutorch.getState() (adcensus.cu:21-29) */
	return v.kind == K_UDATA ? v.ud : nullptr;
}

void lua_pop(lua_State *L, int n)
{
	for (int i = 0; i < n && !L->stack.empty(); i++) L->stack.pop_back();
}

void lua_pushinteger(lua_State *L, lua_Integer n) { Value v; v.kind = K_NUM; v.num = (double)n; L->stack.push_back(v); }

lua_Integer luaL_checkinteger(lua_State *L, int narg)
{
	Value &v = at(L, narg);
	if (v.kind != K_NUM) throw ShimError("bad argument #" + std::to_string(narg) + " (number expected)");
	return (lua_Integer)v.num;
}

lua_Number luaL_checknumber(lua_State *L, int narg)
{
	Value &v = at(L, narg);
	if (v.kind != K_NUM) throw ShimError("bad argument #" + std::to_string(narg) + " (number expected)");
	return v.num;
}

const char *luaL_checkstring(lua_State *L, int narg)
{
	Value &v = at(L, narg);
	if (v.kind != K_STR) throw ShimError("bad argument #" + std::to_string(narg) + " (string expected)");
	return v.str.c_str();
}

int luaL_error(lua_State *, const char *fmt, ...)
{
	va_list ap;
	va_start(ap, fmt);
	std::string s = vformat(fmt, ap);
	va_end(ap);
	throw ShimError(s);
}

void luaL_openlib(lua_State *, const char *libname, const luaL_Reg *l, int)
{
	std::vector<luaL_Reg> &v = registry()[libname];
	for (; l && l->name; l++) {
		bool found = false;
		for (luaL_Reg &r : v)
			if (strcmp(r.name, l->name) == 0) { r.func = l->func; found = true; }  /* re-open: replace */
		if (!found) v.push_back(*l);
	}
}

/* ---- luaT ------------------------------------------------------------ */
void *luaT_checkudata(lua_State *L, int ud, const char *tname)
{
	Value &v = at(L, ud);
	if (v.kind != K_UDATA || v.tname != tname)
		throw ShimError("bad argument #" + std::to_string(ud) + " (" + tname + " expected, got " +
				(v.kind == K_UDATA ? v.tname : std::string("non-tensor")) + ")");
	return v.ud;
}

void luaT_pushudata(lua_State *L, void *udata, const char *tname)
{
	Value v; v.kind = K_UDATA; v.ud = udata; v.tname = tname; L->stack.push_back(v);
}

int luaT_getfieldcheckboolean(lua_State *, int, const char *field) { throw ShimError(std::string("refshim: table field access not supported: ") + field); }
double luaT_getfieldchecknumber(lua_State *, int, const char *field) { throw ShimError(std::string("refshim: table field access not supported: ") + field); }
void *luaT_getfieldcheckudata(lua_State *, int, const char *field, const char *) { throw ShimError(std::string("refshim: table field access not supported: ") + field); }

/* ---- TH / THC -------------------------------------------------------- */
THCudaTensor *THCudaTensor_new(THCState *) { return tensor_new(1, 4); }
void THCudaTensor_resizeAs(THCState *, THCudaTensor *self, THCudaTensor *src) { tensor_resize_as(self, src); }
float *THCudaTensor_data(THCState *, const THCudaTensor *self) { return (float *)self->storage; }
long THCudaTensor_size(THCState *, const THCudaTensor *self, int dim)
{
	if (dim < 0 || dim >= self->nDimension) throw ShimError("THCudaTensor_size: dimension out of range");
	return self->size[dim];
}
long THCudaTensor_nElement(THCState *, const THCudaTensor *self) { return numel(self); }
THCudaTensor *THCudaTensor_newContiguous(THCState *, THCudaTensor *self) { self->refcount++; return self; }
void THCudaTensor_free(THCState *, THCudaTensor *self) { tensor_free(self); }

THFloatTensor *THFloatTensor_new(void) { return (THFloatTensor *)tensor_new(0, 4); }
void THFloatTensor_resizeAs(THFloatTensor *self, THFloatTensor *src) { tensor_resize_as((AnyTensor *)self, (AnyTensor *)src); }
float *THFloatTensor_data(const THFloatTensor *self) { return (float *)self->storage; }
long THFloatTensor_size(const THFloatTensor *self, int dim) { return self->size[dim]; }
long THFloatTensor_nElement(const THFloatTensor *self) { return numel((const AnyTensor *)self); }

double *THDoubleTensor_data(const THDoubleTensor *self) { return (double *)self->storage; }
long THDoubleTensor_size(const THDoubleTensor *self, int dim) { return self->size[dim]; }
long THDoubleTensor_nElement(const THDoubleTensor *self) { return numel((const AnyTensor *)self); }

long *THLongTensor_data(const THLongTensor *self) { return (long *)self->storage; }
long THLongTensor_nElement(const THLongTensor *self) { return numel((const AnyTensor *)self); }

THIntTensor *THIntTensor_new(void) { return (THIntTensor *)tensor_new(0, 4); }
THIntTensor *THIntTensor_newWithSize1d(long size0)
{
	AnyTensor t; memset(&t, 0, sizeof(t)); t.nDimension = 1; t.size[0] = size0;
	AnyTensor *r = tensor_new(0, 4);
	tensor_resize_as(r, &t);
	return (THIntTensor *)r;
}
void THIntTensor_resizeAs(THIntTensor *self, THIntTensor *src) { tensor_resize_as((AnyTensor *)self, (AnyTensor *)src); }
void THIntTensor_zero(THIntTensor *self) { memset(self->storage, 0, (size_t)numel((AnyTensor *)self) * 4); }
int *THIntTensor_data(const THIntTensor *self) { return (int *)self->storage; }

void THError(const char *fmt, ...)
{
	va_list ap;
	va_start(ap, fmt);
	std::string s = vformat(fmt, ap);
	va_end(ap);
	throw ShimError(s);
}

void refshim_argcheck(int cond, int argn, const char *msg)
{
	if (!cond) throw ShimError("bad argument #" + std::to_string(argn) + ": " + msg);
}

/* ---- driver ------------------------------------------------------------ */
lua_State *shim_state_new(void) { return new lua_State(); }

void shim_reset(lua_State *L)
{
	for (Value &v : L->results)
		if (v.kind == K_UDATA) {
			bool wrapped = false;
			for (AnyTensor *w : L->wrappers) wrapped |= ((void *)w == v.ud);
			if (!wrapped) tensor_free((AnyTensor *)v.ud);
		}
	for (AnyTensor *w : L->wrappers) tensor_free(w);
	L->wrappers.clear();
	L->results.clear();
	L->stack.clear();
}

void shim_state_free(lua_State *L)
{
	if (!L) return;
	shim_reset(L);
	delete L;
}

static void push_wrapped(lua_State *L, void *ptr, int nd, const long *sizes, int on_device, int elsize, const char *tname)
{
	AnyTensor *t = tensor_new(on_device, elsize);
	t->nDimension = nd;
	for (int i = 0; i < nd && i < REFSHIM_MAXDIM; i++) t->size[i] = sizes[i];
	t->storage = ptr;
	t->owns = 0;
	L->wrappers.push_back(t);
	luaT_pushudata(L, t, tname);
}

void shim_push_cuda_tensor(lua_State *L, void *devptr, int nd, const long *sizes) { push_wrapped(L, devptr, nd, sizes, 1, 4, "torch.CudaTensor"); }
void shim_push_float_tensor(lua_State *L, void *hostptr, int nd, const long *sizes) { push_wrapped(L, hostptr, nd, sizes, 0, 4, "torch.FloatTensor"); }
void shim_push_number(lua_State *L, double v) { Value x; x.kind = K_NUM; x.num = v; L->stack.push_back(x); }
void shim_push_string(lua_State *L, const char *s) { Value x; x.kind = K_STR; x.str = s; L->stack.push_back(x); }

static lua_CFunction find_fn(const char *lib, const char *fn)
{
	auto it = registry().find(lib);
	if (it == registry().end()) return nullptr;
	for (const luaL_Reg &r : it->second)
		if (strcmp(r.name, fn) == 0) return r.func;
	return nullptr;
}

int shim_call(lua_State *L, const char *lib, const char *fn)
{
	L->error.clear();
	L->results.clear();
	lua_CFunction f = find_fn(lib, fn);
	if (!f) {
		L->error = std::string("attempt to call field '") + fn + "' (a nil value)";
		return -1;
	}
	int nret;
	try {
		nret = f(L);
	} catch (const std::exception &e) {
		L->error = e.what();
		L->stack.clear();
		return -1;
	}
	int n = (int)L->stack.size();
	if (nret > n) nret = n;
	for (int i = n - nret; i < n; i++) L->results.push_back(L->stack[i]);
	L->stack.clear();
	return nret;
}

const char *shim_last_error(lua_State *L) { return L->error.c_str(); }

int shim_result_is_tensor(lua_State *L, int i) { return i >= 0 && i < (int)L->results.size() && L->results[i].kind == K_UDATA; }

void *shim_result_tensor(lua_State *L, int i, int *nd, long *sizes)
{
	if (!shim_result_is_tensor(L, i)) return nullptr;
	AnyTensor *t = (AnyTensor *)L->results[i].ud;
	if (nd) *nd = t->nDimension;
	if (sizes) for (int k = 0; k < t->nDimension; k++) sizes[k] = t->size[k];
	return t->storage;
}

/* device-to-device copy of a returned tensor into caller-owned memory (legacy default stream,
 * the stream every reference kernel launches on) */
int shim_copy_result_to(lua_State *L, int i, void *dst, size_t bytes)
{
	if (!shim_result_is_tensor(L, i)) return -1;
	AnyTensor *t = (AnyTensor *)L->results[i].ud;
	if ((size_t)numel(t) * t->elsize != bytes) return -2;
	return (int)cudaMemcpyAsync(dst, t->storage, bytes, cudaMemcpyDeviceToDevice, 0);
}

double shim_result_number(lua_State *L, int i) { return (i >= 0 && i < (int)L->results.size()) ? L->results[i].num : 0.0; }

int shim_has_function(const char *lib, const char *fn) { return find_fn(lib, fn) != nullptr; }

int shim_num_functions(const char *lib)
{
	auto it = registry().find(lib);
	return it == registry().end() ? 0 : (int)it->second.size();
}

const char *shim_function_name(const char *lib, int i)
{
	auto it = registry().find(lib);
	if (it == registry().end() || i < 0 || i >= (int)it->second.size()) return nullptr;
	return it->second[i].name;
}

void shim_pool_trim(void)
{
	for (auto &kv : pool()) cudaFree(kv.second);
	pool().clear();
}

}  /* extern "C" */

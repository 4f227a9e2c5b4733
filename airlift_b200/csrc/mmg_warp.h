// mmg_warp.h -- the tiny warp interface the warp-cooperative stages are written against, so that tests/emu/ can run the same
// code on the CPU: ballot(f) / sum(f) evaluate f(lane) on every lane and combine; each(f) runs f(lane) on every lane; shfl(v, src)
// gives every lane the value lane src(lane) holds; get(v, lane) broadcasts one lane's value; everything else is computed
// redundantly by all lanes.  Per-lane variables are Var<T>: an array of 32 on the CPU, a register on the device.
#ifndef MMG_WARP_H
#define MMG_WARP_H
#include "mmg_core.h"

MMG_HD int mmg_popc(uint32_t x)
{
#ifdef __CUDA_ARCH__
	return __popc(x);
#else
	return __builtin_popcount(x);
#endif
}
MMG_HD int mmg_ffs(uint32_t x) // 1-based index of the lowest set bit, 0 if none
{
#ifdef __CUDA_ARCH__
	return __ffs((int)x);
#else
	return __builtin_ffs((int)x);
#endif
}


struct WarpEmu { // 32 lanes, one after the other
	template <class T> struct Var { T v[32]; T &operator()(int l) { return v[l]; } const T &operator()(int l) const { return v[l]; } };
	template <class F> unsigned ballot(F &&f) const { unsigned m = 0; for (int l = 0; l < 32; ++l) if (f(l)) m |= 1u << l; return m; }
	template <class F> int sum(F &&f) const { int s = 0; for (int l = 0; l < 32; ++l) s += f(l); return s; }
	template <class F> void each(F &&f) const { for (int l = 0; l < 32; ++l) f(l); }
	template <class F> void one(F &&f) const { f(); }
	template <class F> void regs(F &&f) const { for (int l = 0; l < 32; ++l) f(l); } // per-lane work on Var<> only: nothing goes through memory
	template <class T, class F> Var<T> shfl(const Var<T> &v, F &&src) const { Var<T> o; for (int l = 0; l < 32; ++l) o.v[l] = v.v[src(l) & 31]; return o; }
	template <class T> T get(const Var<T> &v, int lane) const { return v.v[lane & 31]; }
	static void amax(int32_t *p, int32_t v) { if (*p < v) *p = v; }
	static void aadd(int32_t *p, int32_t v) { *p += v; }
};
#ifdef __CUDACC__
struct WarpDev {
	int lane;
	template <class T> struct Var { T v; __device__ T &operator()(int) { return v; } __device__ const T &operator()(int) const { return v; } };
	template <class F> __device__ unsigned ballot(F &&f) const { return __ballot_sync(0xffffffffu, f(lane)); }
	template <class F> __device__ int sum(F &&f) const { return __reduce_add_sync(0xffffffffu, f(lane)); }
	template <class F> __device__ void each(F &&f) const { f(lane); __syncwarp(); }
	template <class F> __device__ void one(F &&f) const { if (lane == 0) f(); __syncwarp(); }
	template <class F> __device__ void regs(F &&f) const { f(lane); }
	template <class T, class F> __device__ Var<T> shfl(const Var<T> &v, F &&src) const { Var<T> o; o.v = __shfl_sync(0xffffffffu, v.v, src(lane)); return o; }
	template <class T> __device__ T get(const Var<T> &v, int src_lane) const { return __shfl_sync(0xffffffffu, v.v, src_lane); }
	static __device__ void amax(int32_t *p, int32_t v) { atomicMax(p, v); }
	static __device__ void aadd(int32_t *p, int32_t v) { atomicAdd(p, v); }
};
#endif
#endif

// klang-b200 — Tier B, first step: a user's `.k` Effect evaluated on the device FROM ITS OWN SOURCE.
//
// The hand-written graphs of kb_graphs.cuh restate 33 reference programs; anything else had no kernel.  tools/kcc.py translates a `.k` file —
// the only change it makes is to mark the program's functions __host__ __device__ and to drop `virtual` — and compiles it with nvcc against THIS
// header, a device-capable subset of the klang.h interface (klang.h:1079-1166 signal arithmetic, 1655-1755 Control, 1796-1830 Dial / Toggle /
// Slider / Button, 2181-2329 the `>>` dataflow protocol for values, 4203-4217 / 4703-4716 the Effect block drivers).  All state is plain data:
// the plugin object is constructed on the host (its constructor fills the controls table exactly as in the reference), copied to the device as
// bytes, and `process()` — the user's body, compiled for the device — runs once per sample:
//   * an Effect with no data members of its own is stateless: thread = sample, a streaming kernel (every sample sees the block's control values);
//   * an Effect with data members runs one lane per instance, frame by frame (Effect::process(buffer), klang.h:4208-4216), its object written back.
// Scope today: programs built from signal / param arithmetic, controls, comparisons, `>>` and libm-free expressions (Gain.k, Pan.k, Clipping.k,
// Functions.k, Mute.k and user edits of them); tanh() is the device restatement of the host's (kb_math.cuh).  The primitives with their own state
// (oscillators, filters, delays, envelopes) are next: their device halves already exist as free functions over POD state (kb_prims.cuh).
// The exported C ABI of a translated program is include/klang_b200_user.h.
#pragma once
#include <cuda_runtime.h>
#include <initializer_list>
#include <string>
#include <vector>
#include <string.h>

#include "kb_math.cuh"

#define KB_KD __host__ __device__ __forceinline__

namespace klang {

struct signal {
	float value;
	KB_KD signal(float v = 0.f) : value(v) {}
	KB_KD operator float() const { return value; }
	KB_KD signal& operator=(float v) { value = v; return *this; }
	KB_KD signal& operator+=(float v) { value += v; return *this; }
	KB_KD signal& operator-=(float v) { value -= v; return *this; }
	KB_KD signal& operator*=(float v) { value *= v; return *this; }
	KB_KD signal& operator/=(float v) { value /= v; return *this; }
	KB_KD signal& operator>>(signal& dst) const { dst.value = value; return dst; }          // `a >> out`            klang.h:2211-2215, 4869-4890
};
typedef signal param;                                                                         // klang.h:1168-1199 (a signal that is passed by value)
KB_KD signal& operator>>(float v, signal& dst) { dst.value = v; return dst; }               // `in * gain >> out`: the expression's value lands in out

struct Control {                                                                              // klang.h:1655-1755 (UI fields reduced to the name)
	const char* name; int type; float min, max, initial; signal value, smoothed;
	KB_KD operator float() const { return value.value; }
	KB_KD operator signal() const { return value; }
	KB_KD float smooth() { smoothed = smoothed.value * 0.999f + (1.f - 0.999f) * value.value; return smoothed; }   // klang.h:1715-1716
	KB_KD Control& set(float x) { value = x < min ? min : (max < x ? max : x); return *this; }                      // std::clamp   klang.h:1725-1728
};
enum { KB_KD_ROTARY = 1, KB_KD_BUTTON, KB_KD_TOGGLE, KB_KD_SLIDER };
inline Control Dial(const char* name, float min = 0.f, float max = 1.f, float initial = 0.f) { return { name, KB_KD_ROTARY, min, max, initial, initial, 0.f }; }
inline Control Slider(const char* name, float min = 0.f, float max = 1.f, float initial = 0.f) { return { name, KB_KD_SLIDER, min, max, initial, initial, 0.f }; }
inline Control Toggle(const char* name, bool initial = false) { return { name, KB_KD_TOGGLE, 0.f, 1.f, initial ? 1.f : 0.f, initial ? 1.f : 0.f, 0.f }; }
inline Control Button(const char* name) { return { name, KB_KD_BUTTON, 0.f, 1.f, 0.f, 0.f, 0.f }; }

struct Controls {                                                                             // klang.h:1893-1925 (Array<Control, 128> reduced to 16)
	Control items[16]; int count;
	Controls() : count(0) { memset(items, 0, sizeof(items)); }
	Controls& operator=(std::initializer_list<Control> list) { count = 0; for (const Control& c : list) if (count < 16) items[count++] = c; return *this; }
	KB_KD Control& operator[](int i) { return items[i]; }
	KB_KD const Control& operator[](int i) const { return items[i]; }
	KB_KD int size() const { return count; }
};

// UI-only objects of the constructors (`hardclip >> graph(-2,2,-2,2)`, klang.h:2536-2840): accepted, ignored
struct Graph { };
inline Graph graph(double = 0, double = 0, double = 0, double = 0) { return Graph(); }
template <class F> inline void operator>>(F, Graph&&) { }
template <class F> inline void operator>>(F, Graph&) { }

struct Effect {                                                                               // klang.h:4203-4217
	signal in, out;
	Controls controls;
	typedef Effect kb_base;
	enum { kb_channels = 1 };
	void prepare() { }
};
namespace Stereo {
	struct signal { klang::signal l, r; KB_KD signal(float a = 0.f, float b = 0.f) : l(a), r(b) {} };
	struct Effect {                                                                           // klang.h:4703-4716
		Stereo::signal in, out;
		Controls controls;
		typedef Stereo::Effect kb_base;
		enum { kb_channels = 2 };
		void prepare() { }
	};
}
namespace stereo = Stereo;
namespace optimised { }
namespace basic { }
namespace minimal { }

// libm as the reference's translation unit sees it (float overloads); the device halves restate the host's functions bit for bit
KB_KD float tanh(float x) {
#ifdef __CUDA_ARCH__
	return kb_tanhf(x);
#else
	return ::tanhf(x);
#endif
}
KB_KD float abs(float x) { return ::fabsf(x); }
KB_KD float sqr(float x) { return x * x; }
KB_KD float cube(float x) { return x * x * x; }

}  // namespace klang

// ------------------------------------------------------------------------------------------------------------- kernels
template <class FX> struct kb_user_traits {
	static constexpr bool stateless = sizeof(FX) == sizeof(typename FX::kb_base);
	static constexpr int channels = FX::kb_channels;
};
template <class FX> KB_KD void kb_user_frame(FX& fx, float* l, float* r) {
	if constexpr (FX::kb_channels == 1) { fx.in = *l; fx.process(); *l = fx.out; }
	else { fx.in.l = *l; fx.in.r = *r; fx.process(); *l = fx.out.l; *r = fx.out.r; }
}
// stateless: thread = sample.  blockIdx.y = instance.
template <class FX> __global__ void kb_user_stream_kernel(const FX* __restrict__ objs, float* __restrict__ io, int n, int stride) {
	__shared__ __align__(16) unsigned char s_raw[sizeof(FX)];            // (raw bytes: the plugin's constructor is host code)
	for (int w = threadIdx.x; w < (int)(sizeof(FX) / 4); w += blockDim.x) reinterpret_cast<unsigned*>(s_raw)[w] = reinterpret_cast<const unsigned*>(objs + blockIdx.y)[w];
	__syncthreads();
	FX fx = *reinterpret_cast<const FX*>(s_raw);
	float* l = io + (size_t)blockIdx.y * FX::kb_channels * stride;
	float* r = l + stride;
	for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) kb_user_frame(fx, l + t, r + t);
}
// stateful: lane = instance, frame by frame; the object (its members are the effect's state) is written back
template <class FX> __global__ void kb_user_seq_kernel(FX* __restrict__ objs, float* __restrict__ io, int n, int stride, int instances) {
	const int inst = blockIdx.x * blockDim.x + threadIdx.x;
	if (inst >= instances) return;
	FX fx = objs[inst];
	float* l = io + (size_t)inst * FX::kb_channels * stride;
	float* r = l + stride;
	for (int t = 0; t < n; t++) kb_user_frame(fx, l + t, r + t);
	objs[inst] = fx;
}

// ------------------------------------------------------------------------------------------------------------- the program's C ABI
struct kb_user_fx_base { virtual ~kb_user_fx_base() {} };
static thread_local std::string kb_user_err;
template <class FX> struct kb_user_bank : kb_user_fx_base {
	int instances = 0, max_block = 0, device = 0;
	std::vector<FX> host;
	FX* d_objs = nullptr; float* d_io = nullptr;
	cudaStream_t stream = nullptr;
	bool dirty = true, host_stale = false;
	static_assert(sizeof(FX) % 4 == 0, "plugin objects are copied word by word");
	int fetch() {
		if (!host_stale) return 0;
		if (cudaMemcpyAsync(host.data(), d_objs, sizeof(FX) * instances, cudaMemcpyDeviceToHost, stream) != cudaSuccess || cudaStreamSynchronize(stream) != cudaSuccess) return -2;
		host_stale = false;
		return 0;
	}
	int process(float* io, int n, unsigned flags) {
		if (n < 0 || n > max_block || !io) { kb_user_err = "kb_user_fx_process: bad argument (n > max_block?)"; return -1; }
		if (n == 0) return 0;
		cudaSetDevice(device);
		// Effect::process(buffer) calls prepare() once per block (klang.h:4209): event-rate code, on the host mirror
		if (fetch()) { kb_user_err = "kb_user_fx_process: state fetch failed"; return -2; }
		for (FX& fx : host) fx.prepare();
		if (cudaMemcpyAsync(d_objs, host.data(), sizeof(FX) * instances, cudaMemcpyHostToDevice, stream) != cudaSuccess) { kb_user_err = "kb_user_fx_process: upload failed"; return -2; }
		cudaStreamSynchronize(stream);                                   // (the mirror is pageable and may change before the copy engine has read it)
		const size_t floats = (size_t)instances * FX::kb_channels * n;
		float* d = io;
		if (!(flags & 1u)) { d = d_io; if (cudaMemcpyAsync(d, io, floats * 4, cudaMemcpyHostToDevice, stream) != cudaSuccess) { kb_user_err = "kb_user_fx_process: H2D failed"; return -2; } }
		if constexpr (kb_user_traits<FX>::stateless) {
			dim3 grid((unsigned)std::max(1, std::min((n + 255) / 256, 148 * 8 / std::min(instances, 148 * 8) + 1)), instances);
			kb_user_stream_kernel<FX><<<grid, 256, 0, stream>>>(d_objs, d, n, n);
		} else {
			kb_user_seq_kernel<FX><<<(instances + 31) / 32, 32, 0, stream>>>(d_objs, d, n, n, instances);
			host_stale = true;
		}
		if (cudaGetLastError() != cudaSuccess) { kb_user_err = "kb_user_fx_process: launch failed"; return -2; }
		if (!(flags & 1u)) {
			if (cudaMemcpyAsync(io, d, floats * 4, cudaMemcpyDeviceToHost, stream) != cudaSuccess || cudaStreamSynchronize(stream) != cudaSuccess) { kb_user_err = "kb_user_fx_process: D2H failed"; return -2; }
		}
		return 0;
	}
	~kb_user_bank() { cudaSetDevice(device); if (stream) cudaStreamSynchronize(stream); cudaFree(d_objs); cudaFree(d_io); if (stream) cudaStreamDestroy(stream); }
};

#define KB_USER_EXPORT(FX, NAME)                                                                                                             \
	extern "C" const char* kb_user_name(void) { return NAME; }                                                                               \
	extern "C" int kb_user_channels(void) { return FX::kb_channels; }                                                                        \
	extern "C" int kb_user_stateless(void) { return kb_user_traits<FX>::stateless ? 1 : 0; }                                                  \
	extern "C" const char* kb_user_last_error(void) { return kb_user_err.c_str(); }                                                           \
	extern "C" int kb_user_num_controls(void) { FX fx; return fx.controls.size(); }                                                           \
	extern "C" void* kb_user_fx_create(int instances, float fs, int max_block, int device) {                                                  \
		(void)fs;                                                                                                                             \
		int ndev = 0;                                                                                                                         \
		if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) { cudaGetLastError(); kb_user_err = "kb_user_fx_create: no such CUDA device (there is no CPU path)"; return nullptr; } \
		if (instances < 1 || instances > 32767 || max_block < 1) { kb_user_err = "kb_user_fx_create: bad argument"; return nullptr; }          \
		kb_user_bank<FX>* b = new kb_user_bank<FX>();                                                                                         \
		b->instances = instances; b->max_block = max_block; b->device = device;                                                               \
		b->host.resize(instances);                                                                                                            \
		bool ok = cudaSetDevice(device) == cudaSuccess && cudaStreamCreateWithFlags(&b->stream, cudaStreamNonBlocking) == cudaSuccess &&       \
		          cudaMalloc(&b->d_objs, sizeof(FX) * instances) == cudaSuccess &&                                                            \
		          cudaMalloc(&b->d_io, sizeof(float) * (size_t)instances * FX::kb_channels * max_block) == cudaSuccess;                        \
		if (!ok) { kb_user_err = "kb_user_fx_create: CUDA allocation failed"; delete b; return nullptr; }                                      \
		return b;                                                                                                                             \
	}                                                                                                                                         \
	extern "C" void kb_user_fx_destroy(void* p) { delete static_cast<kb_user_bank<FX>*>(p); }                                                  \
	extern "C" int kb_user_fx_set_control(void* p, int inst, int idx, float v) {                                                              \
		kb_user_bank<FX>* b = static_cast<kb_user_bank<FX>*>(p);                                                                              \
		if (!b || inst < 0 || inst >= b->instances || idx < 0 || idx >= b->host[0].controls.size()) { kb_user_err = "kb_user_fx_set_control: bad argument"; return -1; } \
		if (b->fetch()) return -2;                                                                                                            \
		b->host[inst].controls[idx].set(v);                                                                                                   \
		return 0;                                                                                                                             \
	}                                                                                                                                         \
	extern "C" int kb_user_fx_get_control(void* p, int inst, int idx, float* v) {                                                             \
		kb_user_bank<FX>* b = static_cast<kb_user_bank<FX>*>(p);                                                                              \
		if (!b || !v || inst < 0 || inst >= b->instances || idx < 0 || idx >= b->host[0].controls.size()) { kb_user_err = "kb_user_fx_get_control: bad argument"; return -1; } \
		if (b->fetch()) return -2;                                                                                                            \
		*v = b->host[inst].controls[idx].value;                                                                                               \
		return 0;                                                                                                                             \
	}                                                                                                                                         \
	extern "C" int kb_user_fx_process(void* p, float* io, int n, unsigned flags) {                                                            \
		kb_user_bank<FX>* b = static_cast<kb_user_bank<FX>*>(p);                                                                              \
		if (!b) { kb_user_err = "kb_user_fx_process: null bank"; return -1; }                                                                  \
		return b->process(io, n, flags);                                                                                                      \
	}

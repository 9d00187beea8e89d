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
#include <type_traits>
#include <utility>
#include <vector>
#include <algorithm>
#include <string.h>

#include "kb_prims.cuh"

#define KB_KD __host__ __device__ __forceinline__

// klang::fs (klang.h:1593-1604) as translated programs see it: kcc rewrites the identifier `fs` to kb_fs(), which reads the bank's sample
// rate from constant memory on the device and from a global on the host (constructors, prepare()).
__constant__ KbFs kb_kd_fs_dev;
static KbFs kb_kd_fs_host = { 44100.f, 44100, 1.f / 44100.f, 2.0f * KB_PI_F * (1.f / 44100.f), 22050.f };

namespace klang {

typedef void event;                                                                            // klang.h:242-243
KB_KD float kb_kd_expf(float x) {                                                             // expf as the reference's translation unit sees it: the host's, restated on the device
#ifdef __CUDA_ARCH__
	return kb_expf(x);
#else
	return ::expf(x);
#endif
}
struct constant {                                                                              // klang.h:93-113: a double, its float and the float of its inverse
	double d; float f; float inv;
	KB_KD constexpr constant(double value) : d(value), f((float)value), inv((float)(1.0 / value)) {}
	KB_KD constexpr operator float() const { return f; }
	KB_KD constexpr float operator^(float x) const { return f; }
};
// pi, ln2, root2 (klang.h:227-233) are namespace-scope objects there; device code cannot name a host object, so kcc rewrites the identifiers
// to these functions (the controls tables of the constructors — host code — use them as well)
KB_KD constexpr constant kb_pi() { return constant(3.1415926535897932384626433832795); }
KB_KD constexpr constant kb_ln2() { return constant(0.6931471805599453094172321214581); }
KB_KD constexpr constant kb_root2() { return constant(1.4142135623730950488016887242097); }

// klang.h:221-224: the first argument's type is the result's — `max(1000, f0 * 1.5)` is an int.  kcc rewrites the program's unqualified
// min( / max( to these (the CUDA headers' own overloads of ::min / ::max would make the call ambiguous)
template <class A, class B> KB_KD A kb_min(A a, B b) { return a < b ? a : (A)b; }
template <class A, class B> KB_KD A kb_max(A a, B b) { return a > b ? a : (A)b; }

struct SampleRate {                                                                            // klang.h:1593-1604
	float f; int i; float inv, w, nyquist; KbFs k;
	KB_KD SampleRate(const KbFs& s) : f(s.f), i(s.i), inv(s.inv), w(s.w), nyquist(s.nyquist), k(s) {}
	KB_KD operator float() const { return f; }
};
KB_KD SampleRate kb_fs() {
#ifdef __CUDA_ARCH__
	return SampleRate(kb_kd_fs_dev);
#else
	return SampleRate(kb_kd_fs_host);
#endif
}

struct signal {                                                                               // klang.h:1062-1166: the operator set as written there,
	float value;                                                                              // so that every expression has the reference's types
	KB_KD constexpr signal(float v = 0.f) : value(v) {}                                      // (constexpr: a program's namespace-scope constants
	KB_KD constexpr signal(double v) : value((float)v) {}                                    //  are also emitted as __device__ objects, which need
	KB_KD constexpr signal(int v) : value((float)v) {}                                       //  constant initialisation)
	KB_KD signal(const constant& c) : value(c.f) {}                                          // klang.h:1067
	KB_KD const signal& operator<<(const signal& input) { value = input.value; return *this; }   // feedback operator   klang.h:1079-1083
	KB_KD signal& operator>>(signal& dst) const { dst.value = value; return dst; }               // `a >> out`          klang.h:1085-1089
	KB_KD signal& operator+=(const signal& x) { value += x.value; return *this; }
	KB_KD signal& operator-=(const signal& x) { value -= x.value; return *this; }
	KB_KD signal& operator*=(const signal& x) { value *= x.value; return *this; }
	KB_KD signal& operator/=(const signal& x) { value /= x.value; return *this; }
	KB_KD signal& operator+=(float x) { value += x; return *this; }
	KB_KD signal& operator-=(float x) { value -= x; return *this; }
	KB_KD signal& operator*=(float x) { value *= x; return *this; }
	KB_KD signal& operator/=(float x) { value /= x; return *this; }
	KB_KD signal& operator+=(double x) { value += (float)x; return *this; }
	KB_KD signal& operator-=(double x) { value -= (float)x; return *this; }
	KB_KD signal& operator*=(double x) { value *= (float)x; return *this; }
	KB_KD signal& operator/=(double x) { value /= (float)x; return *this; }
	KB_KD signal& operator+=(int x) { value += (float)x; return *this; }
	KB_KD signal& operator-=(int x) { value -= (float)x; return *this; }
	KB_KD signal& operator*=(int x) { value *= (float)x; return *this; }
	KB_KD signal& operator/=(int x) { value /= (float)x; return *this; }
	KB_KD signal operator+(float x) const { return value + x; }
	KB_KD signal operator-(float x) const { return value - x; }
	KB_KD signal operator*(float x) const { return value * x; }
	KB_KD signal operator/(float x) const { return value / x; }
	KB_KD signal operator+(double x) const { return value + (float)x; }
	KB_KD signal operator-(double x) const { return value - (float)x; }
	KB_KD signal operator*(double x) const { return value * (float)x; }
	KB_KD signal operator/(double x) const { return value / (float)x; }
	KB_KD signal operator+(int x) const { return value + (float)x; }
	KB_KD signal operator-(int x) const { return value - (float)x; }
	KB_KD signal operator*(int x) const { return value * (float)x; }
	KB_KD signal operator/(int x) const { return value / (float)x; }
	KB_KD operator float() const { return value; }
	KB_KD operator float&() { return value; }
};
typedef signal param;                                                                         // klang.h:1357-1371 (a signal that is passed by value)
KB_KD signal& operator>>(float v, signal& dst) { dst.value = v; return dst; }               // `in * gain >> out`   klang.h:1196-1199

struct Control {                                                                              // klang.h:1655-1755 (UI fields reduced to the name)
	const char* name; int type; float min, max, initial; signal value, smoothed;
	KB_KD operator float() const { return value.value; }
	KB_KD operator signal() const { return value; }
	KB_KD signal operator+(float x) const { return value.value + x; }                        // (Control derives from signal in the reference: its
	KB_KD signal operator-(float x) const { return value.value - x; }                        //  arithmetic is signal's)
	KB_KD signal operator*(float x) const { return value.value * x; }
	KB_KD signal operator/(float x) const { return value.value / x; }
	KB_KD float smooth() { smoothed = smoothed.value * 0.999f + (1.f - 0.999f) * value.value; return smoothed; }   // klang.h:1715-1716
	KB_KD Control& set(float x) { value = x < min ? min : (max < x ? max : x); return *this; }                      // std::clamp   klang.h:1725-1728
};
enum { KB_KD_ROTARY = 1, KB_KD_BUTTON, KB_KD_TOGGLE, KB_KD_SLIDER, KB_KD_MENU, KB_KD_METER };
inline Control Dial(const char* name, float min = 0.f, float max = 1.f, float initial = 0.f) { return { name, KB_KD_ROTARY, min, max, initial, initial, 0.f }; }
inline Control Slider(const char* name, float min = 0.f, float max = 1.f, float initial = 0.f) { return { name, KB_KD_SLIDER, min, max, initial, initial, 0.f }; }
inline Control Toggle(const char* name, bool initial = false) { return { name, KB_KD_TOGGLE, 0.f, 1.f, initial ? 1.f : 0.f, initial ? 1.f : 0.f, 0.f }; }
template <class... Options> inline Control Menu(const char* name, const Options*...) {           // klang.h:1816-1825: value = index of the option, 0 .. count - 1
	return { name, KB_KD_MENU, 0.f, (float)sizeof...(Options) - 1.f, 0.f, 0.f, 0.f };
}
inline Control Button(const char* name) { return { name, KB_KD_BUTTON, 0.f, 1.f, 0.f, 0.f, 0.f }; }
inline Control Meter(const char* name, float min = 0.f, float max = 1.f, float initial = 0.f) { return { name, KB_KD_METER, min, max, initial, initial, 0.f }; }   // klang.h:1838-1841
enum Mode { Peak, RMS, Mean };                                                                // klang.h:88
// A member that stands for one of the plugin's controls (klang.h:1775-1790).  The reference keeps a Control*; an object that travels to the
// device as bytes cannot, so the link is the distance from this member to the control — both live inside the plugin object and move together.
struct ControlMap {
	long long rel;
	KB_KD ControlMap() : rel(0) {}
	KB_KD ControlMap(Control& c) : rel((const char*)&c - (const char*)this) {}
	KB_KD ControlMap(const ControlMap& o) : rel(o.rel ? (const char*)&o + o.rel - (const char*)this : 0) {}
	KB_KD ControlMap& operator=(const ControlMap& o) { rel = o.rel ? (const char*)&o + o.rel - (const char*)this : 0; return *this; }
	KB_KD ControlMap& operator=(Control& c) { rel = (const char*)&c - (const char*)this; return *this; }
	KB_KD Control& control() const { return *(Control*)((char*)this + rel); }
	KB_KD operator signal&() { return control().value; }
	KB_KD operator float() const { return control().value.value; }
	KB_KD signal smooth() { return control().smooth(); }
};
KB_KD signal& operator>>(float v, ControlMap& m) { m.control().value = v; return m.control().value; }   // `gain(level) >> meter`
KB_KD signal& operator>>(float v, Control& c) { c.value = v; return c.value; }

struct ControlGroup {                                                                         // a control, or `{ "Caption", Dial(...), Dial(...) }`: a captioned group whose
	Control items[8]; int count;                                                              // members join the table in order   klang.h:1835-1891
	ControlGroup(const Control& c) : count(1) { items[0] = c; }
	template <class... C> ControlGroup(const char*, C... members) : count(0) { const Control all[] = { members... }; for (const Control& c : all) if (count < 8) items[count++] = c; }
};
struct Controls {                                                                             // klang.h:1893-1925 (Array<Control, 128> reduced to 16)
	Control items[16]; int count;
	Controls() : count(0) { memset(items, 0, sizeof(items)); }
	Controls& operator=(std::initializer_list<ControlGroup> list) {
		count = 0;
		for (const ControlGroup& g : list) for (int k = 0; k < g.count; k++) if (count < 16) items[count++] = g.items[k];
		return *this;
	}
	KB_KD Control& operator[](int i) { return items[i]; }
	KB_KD const Control& operator[](int i) const { return items[i]; }
	KB_KD int size() const { return count; }
};

// UI-only objects of the constructors (`hardclip >> graph(-2,2,-2,2)`, klang.h:2536-2840): accepted, ignored
struct Graph { void clear() { } template <class T> void add(T) { } };                        // (`graph.clear()` / `graph.add(y)` in a note's on(): kcc routes `graph.` to kb_graph().)
inline Graph graph(double = 0, double = 0, double = 0, double = 0) { return Graph(); }
inline Graph kb_graph() { return Graph(); }

// Lookup tables filled at start-up from a function (klang.h:3303-3400): host objects — the examples read them in constructors and on() only.
template <class TYPE> struct Result {                                                         // klang.h:3306-3328
	TYPE* y; int i; TYPE sum;
	Result(TYPE* array, int index) : y(&array[index]), i(index), sum(0) { }
	TYPE& operator[](int index) { return *(y + index); }
	operator TYPE() const { return *y; }
	Result& operator=(const TYPE& in) { *y = in; return *this; }
	TYPE& operator++(int) { i++; return *++y; }
};
#define FUNCTION(type) (void(*)(type, klang::Result<type>&))[](type x, klang::Result<type>& y)
template <class TYPE, int SIZE> struct Table {                                                // klang.h:3330-3400 (Array<TYPE, SIZE> reduced to its items)
	TYPE items[SIZE]; int count;
	Table(TYPE (*function)(TYPE)) : count(0) { for (int x = 0; x < SIZE; x++) items[count++] = function((TYPE)x); }
	Table(void (*function)(TYPE, Result<TYPE>&)) : count(SIZE) {
		Result<TYPE> y(items, 0);
		for (int x = 0; x < SIZE; x++) { function((TYPE)x, y); y.sum += items[x]; y++; }
	}
	Table(std::initializer_list<TYPE> values) : count(0) { for (TYPE v : values) if (count < SIZE) items[count++] = v; }
	TYPE operator[](int index) const { return items[index]; }
};
template <class F> inline void operator>>(F, Graph&&) { }
template <class F> inline void operator>>(F, Graph&) { }

// ------------------------------------------------------------------------------------------------ the dataflow protocol (klang.h:2181-2329, 4869-4890)
// Generic::Output / Generator / Input / Modifier with the reference's rules: reading an object's output (conversion to signal, arithmetic,
// `obj >> dst`) TICKS its process(); feeding an object (`src >> obj`, `obj << src`) stores the value in `in` and runs its input() hook; `obj(args)`
// calls set(args) and yields the object.  The reference dispatches through virtual functions; objects that travel to the device as bytes cannot
// carry a host vptr, so the dispatch here is static (CRTP over the concrete type), which is the same call for every concrete object.
struct kb_output_tag {};
struct kb_input_tag {};
template <class D> struct OutputT : kb_output_tag {
	signal out;
	KB_KD D& kb_self() { return *static_cast<D*>(this); }
	KB_KD operator signal() { kb_self().process(); return out; }                              // processed output        klang.h:2214
	KB_KD operator signal() const { return out; }                                             // last output             klang.h:2215
	KB_KD void reset() { out = 0.f; }
};
template <class S> KB_KD signal kb_read(S& s) { signal v = s; return v; }                     // (an Output lvalue ticks; a value is copied)
template <class S> KB_KD signal kb_read(const S& s) { signal v = s; return v; }
template <class D> struct GeneratorT : OutputT<D> {
	template <class... P> KB_KD D& operator()(P... p) { this->kb_self().set(p...); return this->kb_self(); }   // klang.h:2251-2254
};
template <class D> struct ModifierT : kb_input_tag, OutputT<D> {
	signal in;
	KB_KD void input() {}                                                                     // pre-processing hook     klang.h:2199-2200
	KB_KD void input(const signal& source) { in = source; this->kb_self().input(); }           //                         klang.h:2195
	KB_KD void operator<<(const signal& source) { in = source; this->kb_self().input(); }      // feedback input          klang.h:2194
	KB_KD void process() { this->out = in; }                                                  // pass-through            klang.h:2296
	template <class... P> KB_KD D& operator()(P... p) { this->kb_self().set(p...); return this->kb_self(); }   // klang.h:2299-2302
};
// `source >> destination` (klang.h:4869-4890): an Input receives it through input(); a signal is assigned the (processed) value
template <class S, class D, typename std::enable_if<std::is_base_of<kb_input_tag, D>::value, int>::type = 0>
KB_KD D& operator>>(S&& source, D& destination) { destination.input(kb_read(source)); return destination; }
template <class S, typename std::enable_if<std::is_base_of<kb_output_tag, typename std::remove_reference<S>::type>::value, int>::type = 0>
KB_KD signal& operator>>(S&& source, signal& destination) { destination = kb_read(source); return destination; }
// arithmetic on an Output ticks it once (klang.h:2218-2244); the other operand arrives as float like there
#define KB_KD_OUT_OP(OP)                                                                                                                   \
	template <class D> KB_KD signal operator OP(OutputT<D>& o, float x) { return kb_read(static_cast<D&>(o)) OP x; }                        \
	template <class D> KB_KD signal operator OP(float x, OutputT<D>& o) { return signal(x) OP (float)kb_read(static_cast<D&>(o)); }          \
	template <class D, class E> KB_KD signal operator OP(OutputT<D>& o, OutputT<E>& p) { const signal a = kb_read(static_cast<D&>(o)); return a OP (float)kb_read(static_cast<E&>(p)); }
KB_KD_OUT_OP(+) KB_KD_OUT_OP(-) KB_KD_OUT_OP(*) KB_KD_OUT_OP(/)
#undef KB_KD_OUT_OP

// Function<Args...> (klang.h:2331-2510, 2929-2941): a Modifier that applies a C function to its input and up to two stored arguments.  The
// reference holds the function in a std::function — a host address; kcc reads the function's name from the constructor's initialiser
// (`Shaping() : f(softclip)`) and names it in the member's TYPE (FN::call), so the call is static.  `f(args)` with one argument fewer than the
// function takes stores them and leaves x to the input (`in >> f(distort) >> out`); with all of them it also sets x.
template <class FN, int ARGS> struct FunctionT : ModifierT<FunctionT<FN, ARGS>> {
	float a[3];
	FunctionT() { a[0] = a[1] = a[2] = 0.f; }
	KB_KD float evaluate() const {
		if constexpr (ARGS == 1) return FN::call(this->in.value);
		else if constexpr (ARGS == 2) return FN::call(this->in.value, a[0]);
		else return FN::call(this->in.value, a[0], a[1]);
	}
	KB_KD void process() { this->out = evaluate(); }
	KB_KD FunctionT& operator()(float x0) { if (ARGS == 1) this->in = x0; else a[0] = x0; return *this; }
	KB_KD FunctionT& operator()(float x0, float x1) { if (ARGS == 2) { this->in = x0; a[0] = x1; } else { a[0] = x0; a[1] = x1; } return *this; }
	KB_KD FunctionT& operator()(float x0, float x1, float x2) { this->in = x0; a[0] = x1; a[1] = x2; return *this; }
};

// `x >> debug` (klang.h:3132-3287): kcc rewrites the sink to a temporary of this type; the source is still evaluated (an Output ticks)
struct Debug { };
template <class S> KB_KD signal operator>>(S&& source, Debug&&) { return kb_read(source); }     // (`(in >> debug) >> follower`: the tapped value flows on)
// an Output into a control (`follower >> controls[0]`: a meter): ticks the source   klang.h:2209-2210
template <class S, typename std::enable_if<std::is_base_of<kb_output_tag, typename std::remove_reference<S>::type>::value, int>::type = 0>
KB_KD signal& operator>>(S&& source, Control& destination) { destination.value = kb_read(source); return destination.value; }

// ------------------------------------------------------------------------------------------------ primitives with state (klang.h regions cited per class)
// Thin klang-shaped classes over the POD state and the __host__ __device__ functions of kb_prims.cuh — the functions the hand-written graphs
// are made of, so a translated program and the bound graph of the same `.k` file execute the same arithmetic.
namespace Generators { namespace Fast {
	struct Sine : GeneratorT<Sine>, KbFastSine {                                              // Fast::Sine   klang.h:5135-5172 (`frequency` is a public member there too)
		Sine() { kb_fsine_init(*this); }
		KB_KD void reset() { position = 0u; offset = 0u; }                                    // klang.h:5136-5140 (set(frequency, 0) finds the frequency unchanged)
		KB_KD void set(param f) { kb_fsine_set_f(kb_fs().k, *this, f); }
		KB_KD void set(param f, param phase) { kb_fsine_set_fp(kb_fs().k, *this, f, phase); }
		KB_KD void process() { out = kb_fsine_tick(*this); }
	};
	template <int WAVEFORM, int DUTY_PERCENT> struct OsmT : GeneratorT<OsmT<WAVEFORM, DUTY_PERCENT>> {   // Fast::Saw / Triangle / Square / Pulse   klang.h:5175-5354
		KbOsm o;
		OsmT() { kb_osm_construct(o, WAVEFORM, DUTY_PERCENT / 100.f); }
		KB_KD void set(param f) { kb_osm_set_f(kb_fs().k, o, f); }
		KB_KD void set(param f, param phase) { kb_osm_set_fp(kb_fs().k, o, f, phase); }
		KB_KD void set(param f, param phase, param duty) { kb_osm_set_fpd(kb_fs().k, o, f, phase, duty); }
		KB_KD void process() { this->out = kb_osm_tick(o); }
	};
	typedef OsmT<0, 0> Saw; typedef OsmT<0, 100> Triangle; typedef OsmT<1, 100> Square; typedef OsmT<1, 50> Pulse;
}
namespace Basic {                                                                             // Generators::Basic   klang.h:4897-4944 (float phase, libm sine)
	template <int SHAPE> struct OscT : GeneratorT<OscT<SHAPE>>, KbBasicOsc {
		OscT() { kb_bosc_init(*this); }
		KB_KD void reset() { position = 0.f; }                                                // Oscillator::reset   klang.h:2859
		KB_KD void set(param f) { kb_bosc_set_f(kb_fs().k, *this, f); }
		KB_KD void set(param f, param phase) { kb_bosc_set_fp(kb_fs().k, *this, f, phase); }
		KB_KD void set(param f, param phase, param duty_) { kb_bosc_set_fp(kb_fs().k, *this, f, phase); duty = duty_; }   // Pulse   klang.h:4935-4938
		KB_KD void process() { this->out = SHAPE == 0 ? kb_bosc_sine_tick(*this) : kb_bosc_shape_tick(*this, SHAPE); }
	};
	typedef OscT<0> Sine; typedef OscT<KB_BOSC_SAW> Saw; typedef OscT<KB_BOSC_TRIANGLE> Triangle; typedef OscT<KB_BOSC_SQUARE> Square; typedef OscT<KB_BOSC_PULSE> Pulse;
} }
// Noise (klang.h:4947-4951 Basic, 5357-5366 Fast): one libc rand() per tick from the PROCESS-WIDE stream.  The device continues that stream
// (kb_rand.h: glibc's TYPE_3 generator, jump-ahead, hand-over of the live libc state): before a block the host gives every Noise object the
// generator positioned where the reference's draws for that object begin — objects in the reference's processing order (instance by instance,
// inside a synth note by note over the notes that are not Off), each ticking once per sample — and afterwards advances libc by the draws the
// block consumed.  The objects are found through the registry their constructors fill while the bank builds its host mirror.  Programs with
// several Noise objects per plugin / note interleave their draws sample by sample in declaration order (`skip` draws between two ticks).
struct kb_noise_state { KbRand g; unsigned skip, draws; };
static thread_local std::vector<void*> kb_noise_registry;
template <bool FAST> struct NoiseT : GeneratorT<NoiseT<FAST>> {
	kb_noise_state kb;
	NoiseT() { memset(&kb, 0, sizeof(kb)); kb_noise_registry.push_back(&kb); }
	KB_KD NoiseT(const NoiseT& o) : kb(o.kb) { this->out = o.out; }       // (copies — a vector growing, a lane's private copy — are not new objects)
	KB_KD void process() {
		const uint32_t r = kb_rand_next(kb.g);
		for (unsigned k = 0; k < kb.skip; k++) kb_rand_next(kb.g);
		kb.draws++;
		this->out = FAST ? kb_noise_fast(r) : kb_noise_basic(r);
	}
};
namespace Generators { namespace Fast { typedef NoiseT<true> Noise; } namespace Basic { typedef NoiseT<false> Noise; } }

namespace Filters {
	template <int ORDER = 1> struct IIR;
	template <> struct IIR<1> : ModifierT<IIR<1>> {                                           // optimised first-order IIR   klang.h:5433-5447
		float a, b;
		IIR() : a(1.f), b(0.f) {}
		KB_KD void set(param coeff) { a = coeff; b = 1.f - a; }
		KB_KD void process() { out = in * a + out * b; }
	};
}
namespace Filters { namespace OnePole {
	template <int TYPE> struct FilterT : ModifierT<FilterT<TYPE>> {                           // OnePole::Filter + LPF / HPF   klang.h:5470-5545
		KbOnePole p;
		FilterT() { kb_onepole_construct(p, TYPE); }
		KB_KD void reset() { kb_onepole_reset(p); }
		KB_KD void set(param f) { kb_onepole_set(kb_fs().k, p, f); }
		KB_KD void process() { p.out = this->out; this->out = kb_onepole_tick(p, this->in); }
	};
	typedef FilterT<KB_OP_LPF> LPF; typedef FilterT<KB_OP_HPF> HPF;
} }
namespace Filters { namespace Biquad {
	template <int TYPE> struct FilterT : ModifierT<FilterT<TYPE>> {                           // Biquad::Filter + LPF / HPF   klang.h:5550-5687
		KbBiquad b;
		FilterT() { kb_biquad_construct(b, TYPE); }
		KB_KD void reset() { kb_biquad_reset(b); }
		KB_KD void set(param f) { kb_biquad_set_f(kb_fs().k, b, f); }
		KB_KD void set(param f, param Q) { kb_biquad_set(kb_fs().k, b, f, Q); }
		KB_KD void process() { this->out = kb_biquad_tick(b, this->in); }
	};
	typedef FilterT<KB_BQ_LPF> LPF; typedef FilterT<KB_BQ_HPF> HPF; typedef FilterT<KB_BQ_BPF> BPF; typedef FilterT<KB_BQ_BRF> BRF;
} }
template <int SIZE> struct Delay : ModifierT<Delay<SIZE>> {                                   // Delay<SIZE>   klang.h:3377-3500
	using ModifierT<Delay<SIZE>>::input;
	float ring[SIZE + 1];                                                                     // buffer of SIZE + 1 floats like klang::buffer
	int position; int last_position; float last_fraction; float time;
	Delay() { clear(); time = 1.f; last_position = 0; last_fraction = 0.f; }
	KB_KD void clear() { for (int i = 0; i <= SIZE; i++) ring[i] = 0.f; position = 0; }
	KB_KD void input() {                                                                      // Delay::input    klang.h:3396-3403
		ring[position] = this->in;
		position++;
		if (position == SIZE) position = 0;
	}
	KB_KD signal tap(int delay) const {                                                       // klang.h:3405-3410
		int read = (position - 1) - delay;
		if (read < 0) read += SIZE;
		return ring[read];
	}
	KB_KD signal tap(float delay) const {                                                     // klang.h:3412-3427
		float read = (float)(position - 1) - delay;
		if (read < 0.f) read += SIZE;
		const int i = (int)read;
		const float fraction = read - i;
		const int j = (i + 1) % SIZE;
		return ring[i] + fraction * (ring[j] - ring[i]);
	}
	KB_KD void set(param samples) {                                                           // Delay::set      klang.h:3480-3489
		time = samples < SIZE ? (float)samples : SIZE;
		float read = (float)(position - 1) - time;
		if (read < 0.f) read += SIZE;
		last_position = (int)read;
		last_fraction = read - last_position;
	}
	KB_KD void process() {                                                                    // Delay::process  klang.h:3461-3473
		const int i = last_position, j = (i + 1) % SIZE;
		this->out = ring[i] + last_fraction * (ring[j] - ring[i]);
		last_position = (last_position + 1) % SIZE;
	}
	KB_KD signal operator()(int delay) const { return tap(delay); }                           // klang.h:3603-3621
	KB_KD signal operator()(float delay) const { return tap(delay); }
	KB_KD signal operator()(double delay) const { return tap((float)delay); }
	KB_KD signal operator()(const signal& delay) const { return tap(delay.value); }
};

template <class D> struct OscillatorT : GeneratorT<D> { };                                    // a program's own `struct X : Oscillator` (kcc names the concrete type)

struct Envelope : GeneratorT<Envelope> {                                                      // Envelope   klang.h:3722-4060 over KbEnv (kb_prims.cuh)
	KbEnv e;
	struct Point { float x, y; template <class A, class B> KB_KD Point(A x_, B y_) : x((float)x_), y((float)y_) {} };   // klang.h:3835-3850
	enum Stage { Sustain = KB_ENV_SUSTAIN, Release = KB_ENV_RELEASE, Off = KB_ENV_OFF };
	struct Follower;                                                                          // klang.h:5861-5900 (defined below)
	Envelope() { kb_env_construct(kb_fs().k, e); }
	KB_KD Envelope& operator=(std::initializer_list<Point> points) {                          // klang.h:3884-3896
		float xy[2 * KB_ENV_MAXPTS]; int n = 0;
		for (const Point& p : points) if (n < KB_ENV_MAXPTS) { xy[2 * n] = p.x; xy[2 * n + 1] = p.y; n++; }
		kb_env_set_points(kb_fs().k, e, n, xy);
		return *this;
	}
	KB_KD void setLoop(int start, int end) { kb_env_set_loop(e, start, end); }                // klang.h:3923-3926
	KB_KD void release(float time = 0.f, float level = 0.f) { kb_env_release(kb_fs().k, e, time, level); }   // klang.h:3961-3966
	KB_KD bool finished() const { return e.stage == KB_ENV_OFF; }
	KB_KD bool operator==(Stage st) const { return e.stage == (int)st; }
	KB_KD bool operator!=(Stage st) const { return e.stage != (int)st; }
	KB_KD void process() { out = kb_env_tick(kb_fs().k, e); }                                 // klang.h:4018-4051
	KB_KD signal& operator++(int) { process(); return out; }                                  // klang.h:4013-4016
};
// FM operator (klang.h:4139-4180): an oscillator whose phase offset is its input, scaled by its own envelope and an amplitude.  The reference
// derives from the oscillator and overrides process() virtually; here the operator re-derives the dataflow base with its own type (static
// dispatch) and carries Fast::Sine's state — the only oscillator the examples put in an Operator.  Restates kb_fm_op_tick (kb_graphs.cuh).
template <class OSCILLATOR> struct Operator;
template <> struct Operator<Generators::Fast::Sine> : GeneratorT<Operator<Generators::Fast::Sine>>, KbFastSine, kb_input_tag {
	Envelope env;
	signal amp, in;
	Operator() : amp(1.f) { kb_fsine_init(*this); }
	KB_KD void reset() { position = 0u; offset = 0u; }
	KB_KD void set(param f) { kb_fsine_set_f(kb_fs().k, *this, f); }
	KB_KD void set(param f, param phase) { kb_fsine_set_fp(kb_fs().k, *this, f, phase); }
	KB_KD void input(const signal& source) { in = source; }                                   // Input::input (no hook)
	KB_KD void operator<<(const signal& source) { in = source; }
	KB_KD Operator& operator=(std::initializer_list<Envelope::Point> points) { env = points; return *this; }   // klang.h:4149-4152
	KB_KD Operator& operator*(signal a) { amp = a; return *this; }                            // klang.h:4159-4162 (hides the Output arithmetic, as there)
	KB_KD void process() {                                                                    // klang.h:4164-4168: set(+in) = Sine::set(relative), process, out *= env++ * amp
		offset = kb_phase_from_radians(in.value * KB_TWO_PI_F);
		this->out = kb_fsine_tick(*this);
		this->out *= (float)(env++) * amp.value;
	}
	KB_KD Operator& operator>>(Operator& carrier) { carrier << kb_read(*this); return carrier; }   // klang.h:4170-4173: `carrier << *this` ticks this operator
};

struct ADSR : Envelope {                                                                      // ADSR   klang.h:4063-4138
	ADSR() { kb_adsr_construct(kb_fs().k, e); }
	KB_KD void set(param A, param D, param S, param R) { kb_adsr_set(kb_fs().k, e, A, D, S, R); }
	KB_KD ADSR& operator()(param A, param D, param S, param R) { set(A, D, S, R); return *this; }
	KB_KD void release() { kb_adsr_release(kb_fs().k, e); }
	using Envelope::release;
};

// Envelope follower (klang.h:5861-5900): |in| or in^2 through an attack / release one-pole, rms takes the root.  The reference selects peak() /
// rms() through a member-function pointer — a host address; here the mode is a flag.  expf is the host's on the host and its restatement on the
// device (kb_math.cuh), sqrt / abs bind to sqrtf / fabsf.
struct Envelope::Follower : ModifierT<Envelope::Follower> {
	signal attack, release, A, R, ar;                                                         // AR: attack = release = 0, A = R = 1, its own out
	int rms_mode;
	Follower() : attack(0.f), release(0.f), A(1.f), R(1.f), ar(0.f), rms_mode(1) { set(0.01f, 0.1f); }
	KB_KD void set(param attack_, param release_) {                                           // AR::set   klang.h:5871-5878
		if (attack.value != attack_.value || release.value != release_.value) {
			attack = attack_; release = release_;
			A = 1.f - (attack_.value == 0.f ? 0.f : kb_kd_expf(-1.0f / (kb_fs().f * attack_.value)));
			R = 1.f - (release_.value == 0.f ? 0.f : kb_kd_expf(-1.0f / (kb_fs().f * release_.value)));
		}
	}
	KB_KD Follower& operator=(Mode mode) { rms_mode = mode == RMS ? 1 : 0; return *this; }
	KB_KD void process() {                                                                    // peak(): abs(in) >> ar >> out;  rms(): (in * in) >> ar >> sqrt >> out
		const float x = rms_mode ? in.value * in.value : ::fabsf(in.value);
		const float smoothing = x > ar.value ? A.value : R.value;                             // AR::process   klang.h:5880-5883
		ar = ar.value + smoothing * (x - ar.value);
		out = rms_mode ? ::sqrtf(ar.value) : ar.value;
	}
};

// host-only pieces of Note::on(): libc rand() and the host libm, exactly where the reference calls them (a device call would be a bug: trap)
template <class T> KB_KD T random(const T min, const T max) {                                 // klang.h:236-237
#ifdef __CUDA_ARCH__
	__trap(); return min;
#else
	return rand() * ((max - min) / (T)RAND_MAX) + min;
#endif
}
KB_KD float power(float base, float e) {                                                      // klang.h:187-218, as Pitch -> Frequency uses it
#ifdef __CUDA_ARCH__
	// base 10 goes through expf in the reference (klang.h:206) and expf is restated for the device (kb_expf); a general powf is not
	if (base == 10.f) return kb_expf(e * (float)2.3025850929940456840179914546843642076011014886287729760333279009);
	else if (e == 0.f) return 1.f;
	else if (e == 1.f) return base;
	else if (e == 2.f) return base * base;
	else if (e == 3.f) return base * base * base;
	__trap(); return base * e;
#else
	if (base == 10.f) return (float)::expf(e * (float)2.3025850929940456840179914546843642076011014886287729760333279009);
	else if (e == 0.f) return 1.f;
	else if (e == 1.f) return base;
	else if (e == 2.f) return base * base;
	else if (e == 3.f) return base * base * base;
	return ::powf(base, e);
#endif
}
struct Amplitude : signal { using signal::signal; KB_KD Amplitude(const signal& s) : signal(s) {} };
typedef Amplitude Velocity;
struct Frequency : signal { using signal::signal; KB_KD constexpr Frequency(const signal& s) : signal(s) {} };   // klang.h:1580-1590 (a param with a unit)
struct dB : signal {                                                                          // klang.h:1609-1621: `x -> Amplitude` (a thread-local there; here
	using signal::signal;                                                                     //  operator-> hands out a value that carries it, so a const dB works)
	struct Conversion { signal Amplitude; KB_KD const Conversion* operator->() const { return this; } };
	KB_KD constexpr dB(const signal& s) : signal(s) {}
	KB_KD dB(const Control& c) : signal(c.value) {}
	KB_KD Conversion operator->() const { Conversion c; c.Amplitude = power(10.f, value * 0.05f); return c; }
};
struct Pitch : signal {                                                                       // klang.h:1551-1578 (Frequency: a member here, a thread-local there)
	using signal::signal;
	signal Frequency;
	KB_KD Pitch(const signal& s) : signal(s) {}
	KB_KD const Pitch* operator->() { Frequency = 440.f * power(2.f, (value - 69.f) / 12.f); return this; }
	template <class T> KB_KD Pitch operator+(T in) const { return Pitch(value + in); }
	template <class T> KB_KD Pitch operator-(T in) const { return Pitch(value - in); }
	template <class T> KB_KD Pitch operator*(T in) const { return Pitch(value * in); }
	template <class T> KB_KD Pitch operator/(T in) const { return Pitch(value / in); }
};

struct Preset { const char* name; float values[16]; int count;                                // klang.h:1940-1981
	Preset() : name(nullptr), count(0) { memset(values, 0, sizeof(values)); }
	Preset(const char* n, std::initializer_list<double> v) : name(n), count(0) { memset(values, 0, sizeof(values)); for (double x : v) if (count < 16) values[count++] = (float)x; } };
struct Presets { Preset items[16]; int count; Presets() : count(0) {}
	Presets& operator=(std::initializer_list<Preset> list) { count = 0; for (const Preset& p : list) if (count < 16) items[count++] = p; return *this; } };

struct Effect {                                                                               // klang.h:4203-4217
	signal in, out;
	Controls controls;
	Presets presets;
	typedef Effect kb_base;
	enum { kb_channels = 1 };
	void prepare() { }
};

// ---- Note / Synth (klang.h:4220-4304, 4376-4467): the note object carries the program's members and travels as bytes; on() / off() run on the
// host mirror (NoteBase::start / release), process() per sample on the device, one lane per voice.
struct NoteControls {                                                                         // NoteBase::Controls: a view of the synth's table   klang.h:4224-4234
	Controls* kb_table;
	KB_KD Control& operator[](int i) { return (*kb_table)[i]; }
	KB_KD const Control& operator[](int i) const { return (*kb_table)[i]; }
	KB_KD unsigned int size() const { return kb_table ? (unsigned)kb_table->size() : 0u; }
};
struct Note {
	signal out;
	Pitch pitch; Velocity velocity;
	NoteControls controls;
	void* kb_synth;
	enum Stage { Onset, Sustain, Release, Off };
	int stage;
	Note() : kb_synth(nullptr), stage(Off) { controls.kb_table = nullptr; }
	KB_KD void on(Pitch, Velocity) { }                                                        // klang.h:4237-4238 (defaults a program may replace)
	KB_KD void off(Velocity = 0) { stage = Off; }
	KB_KD void prepare() { }
	KB_KD bool stop(Velocity = 0) { stage = Off; return true; }                               // klang.h:4277-4280
	KB_KD bool finished() const { return stage == Off; }
	template <class S = void> KB_KD S* getSynth() { return static_cast<S*>(kb_synth); }
};
struct NotesDecl { int count; NotesDecl() : count(0) {} template <class N> void add(int n) { count += n; } };   // Notes::add<NOTE>(n)   klang.h:4325-4331
struct Synth {                                                                                // klang.h:4376-4467 (post-processing: Effect::process over the mix)
	signal in, out;
	Controls controls;
	Presets presets;
	NotesDecl notes;
	typedef Synth kb_base;
	void prepare() { }
	KB_KD void process() { out = in; }
};
namespace Stereo {
	struct signal {                                                                           // Stereo::signal (frame)   klang.h:4484-4560: channel-wise arithmetic
		klang::signal l, r;
		KB_KD signal(float a = 0.f, float b = 0.f) : l(a), r(b) {}
		KB_KD signal operator+(const signal& x) const { return signal(l.value + x.l.value, r.value + x.r.value); }
		KB_KD signal operator-(const signal& x) const { return signal(l.value - x.l.value, r.value - x.r.value); }
		KB_KD signal operator*(const signal& x) const { return signal(l.value * x.l.value, r.value * x.r.value); }
		KB_KD signal operator/(const signal& x) const { return signal(l.value / x.l.value, r.value / x.r.value); }
		KB_KD signal operator+(float x) const { return signal(l.value + x, r.value + x); }
		KB_KD signal operator-(float x) const { return signal(l.value - x, r.value - x); }
		KB_KD signal operator*(float x) const { return signal(l.value * x, r.value * x); }
		KB_KD signal operator/(float x) const { return signal(l.value / x, r.value / x); }
		KB_KD klang::signal& operator[](int index) { return index ? r : l; }                   // signals<2>::operator[]   klang.h:1218-1220
		KB_KD const klang::signal& operator[](int index) const { return index ? r : l; }
	};
	template <int SIZE> struct Delay : kb_input_tag {                                         // Stereo::Delay = Bank<klang::Delay<SIZE>, 2>   klang.h:4645-4699
		klang::Delay<SIZE> items[2];
		Stereo::signal in, out;
		KB_KD void clear() { items[0].clear(); items[1].clear(); }
		KB_KD void input(const Stereo::signal& source) { in = source; items[0].input(source.l); items[1].input(source.r); }   // Bank::input   klang.h:2917-2920
		KB_KD void operator<<(const Stereo::signal& source) { input(source); }
		KB_KD Stereo::signal tap(float delay) const {                                         // both channels at the LEFT line's position   klang.h:4668-4681
			float read = (float)(items[0].position - 1) - delay;
			if (read < 0.f) read += SIZE;
			const float f = floorf(read);
			delay = read - f;
			const int i = (int)read, j = (i == (SIZE - 1)) ? 0 : (i + 1);
			return Stereo::signal(items[0].ring[i] * (1.f - delay) + items[0].ring[j] * delay, items[1].ring[i] * (1.f - delay) + items[1].ring[j] * delay);
		}
		KB_KD Stereo::signal operator()(const Stereo::signal& delay) const { return Stereo::signal(items[0].tap(delay.l.value), items[1].tap(delay.r.value)); }   // klang.h:4687-4696
		KB_KD Stereo::signal operator()(float delay) const { return tap(delay); }
	};
	template <class S> KB_KD Stereo::signal kb_read(const S& s) { return s; }
	struct Effect {                                                                           // klang.h:4703-4716
		Stereo::signal in, out;
		Controls controls;
		Presets presets;
		typedef Stereo::Effect kb_base;
		enum { kb_channels = 2 };
		void prepare() { }
	};
}
namespace stereo = Stereo;
namespace optimised { using namespace Generators::Fast; using namespace Filters; using namespace Filters::Biquad; }   // klang.h:6145-6152
namespace basic { using namespace Generators::Basic; using namespace Filters; using namespace Filters::Biquad; }
namespace minimal { }

// libm as the reference's translation unit sees it (float overloads); the device halves restate the host's functions bit for bit
KB_KD float tanh(float x) {
#ifdef __CUDA_ARCH__
	return kb_tanhf(x);
#else
	return ::tanhf(x);
#endif
}
KB_KD float exp(float x) {                                                                    // (`exp(float)` binds to expf)
#ifdef __CUDA_ARCH__
	return kb_expf(x);
#else
	return ::expf(x);
#endif
}
KB_KD float kb_tanh(float x) { return tanh(x); }                                              // kcc rewrites a program's unqualified tanh( / exp( to these:
KB_KD float kb_exp(float x) { return exp(x); }
KB_KD float kb_abs(float x) { return ::fabsf(x); }                                             // (abs binds to fabsf for a float)
KB_KD int kb_abs(int x) { return x < 0 ? -x : x; }                                                //  with a plain float argument ::tanh(float) of <cmath> would tie
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
	float* l = io + (size_t)inst * FX::kb_channels * stride;
	float* r = l + stride;
	if constexpr (sizeof(FX) <= 4096) {                                  // small state: a private copy the compiler can keep in registers
		FX fx = objs[inst];
		for (int t = 0; t < n; t++) kb_user_frame(fx, l + t, r + t);
		objs[inst] = fx;
	} else {                                                             // objects that hold delay lines stay where they are, in HBM
		FX& fx = objs[inst];
		for (int t = 0; t < n; t++) kb_user_frame(fx, l + t, r + t);
	}
}

// ------------------------------------------------------------------------------------------------------------- the program's C ABI
struct kb_user_fx_base { virtual ~kb_user_fx_base() {} };
static thread_local std::string kb_user_err;
template <class FX> struct kb_user_bank : kb_user_fx_base {
	int instances = 0, max_block = 0, device = 0; KbFs fs;
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
	std::vector<size_t> noise_off;                                       // byte offsets of the Noise states inside the host mirror, in address order
	void find_noise() {
		const char* lo = reinterpret_cast<const char*>(host.data()); const char* hi = lo + sizeof(FX) * host.size();
		for (void* p : klang::kb_noise_registry) if ((const char*)p >= lo && (const char*)p < hi) noise_off.push_back((size_t)((const char*)p - lo));
		std::sort(noise_off.begin(), noise_off.end());
		klang::kb_noise_registry.clear();
	}
	// Noise: hand every object the libc stream at the position of its first draw of this block (instance by instance; inside an instance the
	// objects tick once per sample in declaration order), then advance libc by what the block consumes
	int place_noise(int n) {
		if (noise_off.empty()) return 0;
		const size_t per = noise_off.size() / (size_t)instances;
		KbRand base;
		if (!kb_rand_capture(base)) { kb_user_err = "libc rand() is not running its default (TYPE_3) generator"; return -1; }
		for (size_t k = 0; k < noise_off.size(); k++) {
			klang::kb_noise_state st; st.g = base; st.skip = (unsigned)per - 1u; st.draws = 0u;
			kb_rand_jump(st.g, (unsigned long long)(k / per) * (unsigned long long)n * per + (k % per));
			if (cudaMemcpyAsync(reinterpret_cast<char*>(d_objs) + noise_off[k], &st, sizeof(st), cudaMemcpyHostToDevice, stream) != cudaSuccess) return -2;
		}
		if (cudaStreamSynchronize(stream) != cudaSuccess) return -2;
		kb_rand_jump(base, (unsigned long long)instances * (unsigned long long)n * per);
		return kb_rand_commit(base) ? 0 : -1;
	}
	int process(float* io, int n, unsigned flags) {
		if (n < 0 || n > max_block || !io) { kb_user_err = "kb_user_fx_process: bad argument (n > max_block?)"; return -1; }
		if (n == 0) return 0;
		cudaSetDevice(device);
		// Effect::process(buffer) calls prepare() once per block (klang.h:4209): event-rate code, on the host mirror — a round trip of the objects
		// that programs WITHOUT a prepare() of their own are spared (their objects travel only when a control was set)
		constexpr bool has_prepare = !std::is_same<decltype(&FX::prepare), void (FX::kb_base::*)()>::value;
		if (has_prepare || dirty) {
			if (fetch()) { kb_user_err = "kb_user_fx_process: state fetch failed"; return -2; }
			kb_kd_fs_host = fs;
			for (FX& fx : host) fx.prepare();
			if (cudaMemcpyAsync(d_objs, host.data(), sizeof(FX) * instances, cudaMemcpyHostToDevice, stream) != cudaSuccess) { kb_user_err = "kb_user_fx_process: upload failed"; return -2; }
			cudaStreamSynchronize(stream);                               // (the mirror is pageable and may change before the copy engine has read it)
			dirty = false;
		}
		if (int rc = place_noise(n)) { if (kb_user_err.empty()) kb_user_err = "kb_user_fx_process: noise placement failed"; return rc; }
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
	extern "C" int kb_user_kind(void) { return 0; }                                                                                          \
	extern "C" int kb_user_channels(void) { return FX::kb_channels; }                                                                        \
	extern "C" int kb_user_stateless(void) { return kb_user_traits<FX>::stateless ? 1 : 0; }                                                  \
	extern "C" const char* kb_user_last_error(void) { return kb_user_err.c_str(); }                                                           \
	extern "C" int kb_user_num_controls(void) { std::vector<FX> one(1); return one[0].controls.size(); }                                                        \
	extern "C" void* kb_user_fx_create(int instances, float fs, int max_block, int device) {                                                  \
		if (!(fs > 0.f)) { kb_user_err = "kb_user_fx_create: bad sample rate"; return nullptr; }                                              \
		int ndev = 0;                                                                                                                         \
		if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) { cudaGetLastError(); kb_user_err = "kb_user_fx_create: no such CUDA device (there is no CPU path)"; return nullptr; } \
		if (instances < 1 || instances > 32767 || max_block < 1) { kb_user_err = "kb_user_fx_create: bad argument"; return nullptr; }          \
		kb_user_bank<FX>* b = new kb_user_bank<FX>();                                                                                         \
		b->instances = instances; b->max_block = max_block; b->device = device; b->fs = kb_make_fs(fs);                                        \
		kb_kd_fs_host = b->fs;                                            /* constructors see klang::fs */                                     \
		klang::kb_noise_registry.clear();                                                                                                     \
		b->host.resize(instances);                                                                                                            \
		b->find_noise();                                                                                                                      \
		bool ok = cudaSetDevice(device) == cudaSuccess && cudaStreamCreateWithFlags(&b->stream, cudaStreamNonBlocking) == cudaSuccess &&       \
		          cudaMemcpyToSymbol(kb_kd_fs_dev, &b->fs, sizeof(KbFs)) == cudaSuccess &&                                                     \
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
		b->host[inst].controls[idx].set(v); b->dirty = true;                                                                                  \
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

// ============================================================================================================= translated synths
// Lane = voice: Note::process(buffer) (klang.h:4295-4303) — prepare(), then process() per sample into the voice's stream — for every note
// that was not Off when the block began.  Objects that hold delay lines stay in HBM, small ones are copied to the lane.
template <class NOTE> __global__ void kb_user_note_kernel(NOTE* __restrict__ notes, klang::Controls* __restrict__ controls, int* __restrict__ active,
                                                          float* __restrict__ streams, int n, int voices, int total) {
	const int v = blockIdx.x * blockDim.x + threadIdx.x;
	if (v >= total) return;
	float* o = streams + (size_t)v * n;
	const int act = notes[v].stage != klang::Note::Off;
	active[v] = act;
	if (!act) { for (int t = 0; t < n; t++) o[t] = 0.f; return; }
	auto run = [&](NOTE& nt) {
		nt.controls.kb_table = controls + v / voices;
		nt.prepare();
		for (int t = 0; t < n; t++) { nt.process(); o[t] = nt.out; }
	};
	if constexpr (sizeof(NOTE) <= 4096) { NOTE nt = notes[v]; run(nt); memcpy((void*)&notes[v], (const void*)&nt, sizeof(NOTE)); } else run(notes[v]);   // (bytes: a note with a const member has no operator=)
}
// Synth::process voice loop for a mono synth (klang.h:4450-4456): every active note ASSIGNS the block, so it holds the last active note's
// stream; thread = (instance, sample)
__global__ void kb_user_mono_mix_kernel(const float* __restrict__ streams, const int* __restrict__ active, float* __restrict__ out, int n, int voices) {
	const int inst = blockIdx.y, t = blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= n) return;
	int last = -1;
	for (int v = voices - 1; v >= 0 && last < 0; v--) if (active[inst * voices + v]) last = v;
	out[(size_t)inst * n + t] = last >= 0 ? streams[((size_t)inst * voices + last) * n + t] : 0.f;
}
// the synth's own process() over the mix (Effect::process(buffer), klang.h:4208-4216, called at klang.h:4459): lane = instance
template <class SYN> __global__ void kb_user_post_kernel(SYN* __restrict__ syns, float* __restrict__ out, int n, int instances) {
	const int inst = blockIdx.x * blockDim.x + threadIdx.x;
	if (inst >= instances) return;
	SYN& syn = syns[inst];
	float* o = out + (size_t)inst * n;
	for (int t = 0; t < n; t++) { syn.in = o[t]; syn.process(); o[t] = syn.out; }
}

template <class SYN, class NOTE> struct kb_user_synth_bank : kb_user_fx_base {
	int instances = 0, voices = 0, max_block = 0, device = 0; KbFs fs;
	std::vector<SYN> syn; std::vector<NOTE> notes;                       // host mirrors: the plugin objects (controls) and every note
	std::vector<unsigned> noteOns, noteStart;                            // Notes::noteOns / noteStart   klang.h:4333-4334
	std::vector<unsigned char> dirty_note; bool notes_stale = false, ctl_dirty = true, syn_dirty = true;
	SYN* d_syn = nullptr; NOTE* d_notes = nullptr; klang::Controls* d_ctl = nullptr; int* d_active = nullptr;
	float *d_streams = nullptr, *d_out = nullptr;
	cudaStream_t stream = nullptr;
	static constexpr bool has_post = !std::is_same<decltype(&SYN::process), void (klang::Synth::*)()>::value;
	int total() const { return instances * voices; }
	NOTE& note(int inst, int v) { return notes[(size_t)inst * voices + v]; }
	int fetch() {                                                        // the device has evolved the notes: events need their current state
		if (!notes_stale) return 0;
		if (cudaMemcpyAsync(notes.data(), d_notes, sizeof(NOTE) * total(), cudaMemcpyDeviceToHost, stream) != cudaSuccess || cudaStreamSynchronize(stream) != cudaSuccess) return -2;
		notes_stale = false;
		return 0;
	}
	void bind(int inst, int v) { NOTE& nt = note(inst, v); nt.controls.kb_table = &syn[inst].controls; nt.kb_synth = &syn[inst]; kb_kd_fs_host = fs; }
	// NoteBase::start / release (klang.h:4257-4275) on the host mirror
	int start(int inst, int v, float pitch, float velocity) {
		if (fetch()) return -2;
		bind(inst, v);
		NOTE& nt = note(inst, v);
		nt.stage = klang::Note::Onset; nt.pitch = klang::Pitch(pitch); nt.velocity = klang::Velocity(velocity);
		nt.on(nt.pitch, nt.velocity);
		nt.stage = klang::Note::Sustain;
		dirty_note[(size_t)inst * voices + v] = 1;
		return 0;
	}
	int release(int inst, int v, float velocity) {
		if (fetch()) return -2;
		NOTE& nt = note(inst, v);
		if (nt.stage == klang::Note::Off) return 0;
		if (nt.stage != klang::Note::Release) { bind(inst, v); nt.stage = klang::Note::Release; nt.off(klang::Velocity(velocity)); dirty_note[(size_t)inst * voices + v] = 1; }
		return 0;
	}
	int assign(int inst) {                                               // Notes::assign   klang.h:4336-4372
		unsigned* ns = noteStart.data() + (size_t)inst * voices;
		for (int i = 0; i < voices; i++) if (note(inst, i).stage == klang::Note::Off) { ns[i] = noteOns[inst]++; return i; }
		int oldest = -1; unsigned oldest_start = 0;
		for (int i = 0; i < voices; i++) if (note(inst, i).stage == klang::Note::Release && (oldest == -1 || ns[i] < oldest_start)) { oldest = i; oldest_start = ns[i]; }
		if (oldest != -1) { ns[oldest] = noteOns[inst]++; return oldest; }
		for (int i = 0; i < voices; i++) if (oldest == -1 || ns[i] < oldest_start) { oldest = i; oldest_start = ns[i]; }
		ns[oldest] = noteOns[inst]++;
		return oldest;
	}
	std::vector<size_t> noise_off;                                       // Noise states inside the notes, in address order (note-major)
	void find_noise() {
		const char* lo = reinterpret_cast<const char*>(notes.data()); const char* hi = lo + sizeof(NOTE) * notes.size();
		for (void* p : klang::kb_noise_registry) if ((const char*)p >= lo && (const char*)p < hi) noise_off.push_back((size_t)((const char*)p - lo));
		std::sort(noise_off.begin(), noise_off.end());
		klang::kb_noise_registry.clear();
	}
	// Synth::process renders note after note (klang.h:4450-4456): the notes that are not Off draw n values each, in note order, instance by instance
	int place_noise(int n) {
		if (noise_off.empty()) return 0;
		const size_t per = noise_off.size() / notes.size();
		if (notes_stale) {                                               // only the stages are needed: 4 bytes per note
			std::vector<int> st(notes.size());
			const size_t stage_off = (size_t)(reinterpret_cast<const char*>(&static_cast<const klang::Note&>(notes[0]).stage) - reinterpret_cast<const char*>(&notes[0]));
			if (cudaMemcpy2DAsync(st.data(), sizeof(int), reinterpret_cast<const char*>(d_notes) + stage_off, sizeof(NOTE), sizeof(int), notes.size(), cudaMemcpyDeviceToHost, stream) != cudaSuccess ||
			    cudaStreamSynchronize(stream) != cudaSuccess) return -2;
			for (size_t k = 0; k < notes.size(); k++) if (!dirty_note[k]) notes[k].stage = st[k];
		}
		KbRand base;
		if (!kb_rand_capture(base)) { kb_user_err = "libc rand() is not running its default (TYPE_3) generator"; return -1; }
		unsigned long long at = 0;
		for (size_t v = 0; v < notes.size(); v++) {
			if (notes[v].stage == klang::Note::Off) continue;
			for (size_t j = 0; j < per; j++) {
				klang::kb_noise_state st; st.g = base; st.skip = (unsigned)per - 1u; st.draws = 0u;
				kb_rand_jump(st.g, at + j);
				if (cudaMemcpyAsync(reinterpret_cast<char*>(d_notes) + noise_off[v * per + j], &st, sizeof(st), cudaMemcpyHostToDevice, stream) != cudaSuccess) return -2;
			}
			at += (unsigned long long)n * per;
		}
		if (cudaStreamSynchronize(stream) != cudaSuccess) return -2;
		kb_rand_jump(base, at);
		return kb_rand_commit(base) ? 0 : -1;
	}
	int process(float* out, int n, unsigned flags) {
		if (n < 0 || n > max_block || !out) { kb_user_err = "kb_user_synth_process: bad argument (n > max_block?)"; return -1; }
		if (n == 0) return 0;
		cudaSetDevice(device);
		const bool dev = flags & 1u, per_voice = flags & 2u;
		bool ok = true;
		for (int k = 0; k < total() && ok; k++) if (dirty_note[k]) { ok = cudaMemcpyAsync(d_notes + k, &notes[k], sizeof(NOTE), cudaMemcpyHostToDevice, stream) == cudaSuccess; dirty_note[k] = 0; }
		if (ctl_dirty) {
			std::vector<klang::Controls> c(instances);
			for (int i = 0; i < instances; i++) c[i] = syn[i].controls;
			ok = ok && cudaMemcpyAsync(d_ctl, c.data(), sizeof(klang::Controls) * instances, cudaMemcpyHostToDevice, stream) == cudaSuccess;
			ok = ok && cudaStreamSynchronize(stream) == cudaSuccess;
			ctl_dirty = false;
		}
		if (has_post && syn_dirty) {
			// (the synth object's controls are refreshed; its other members are the post-processing state the device evolves)
			for (int i = 0; i < instances && ok; i++) ok = cudaMemcpyAsync(&d_syn[i].controls, &syn[i].controls, sizeof(klang::Controls), cudaMemcpyHostToDevice, stream) == cudaSuccess;
			syn_dirty = false;
		}
		if (!ok || cudaStreamSynchronize(stream) != cudaSuccess) { kb_user_err = "kb_user_synth_process: upload failed"; return -2; }
		if (int rc = place_noise(n)) { if (kb_user_err.empty()) kb_user_err = "kb_user_synth_process: noise placement failed"; return rc; }
		float* d_streams_dst = (per_voice && dev) ? out : d_streams;
		kb_user_note_kernel<NOTE><<<(total() + 31) / 32, 32, 0, stream>>>(d_notes, d_ctl, d_active, d_streams_dst, n, voices, total());
		notes_stale = true;
		float* d_result = d_streams_dst;
		size_t floats = (size_t)total() * n;
		if (!per_voice) {
			d_result = dev ? out : d_out;
			floats = (size_t)instances * n;
			dim3 grid((n + 255) / 256, instances);
			kb_user_mono_mix_kernel<<<grid, 256, 0, stream>>>(d_streams, d_active, d_result, n, voices);
			if constexpr (has_post) kb_user_post_kernel<SYN><<<(instances + 31) / 32, 32, 0, stream>>>(d_syn, d_result, n, instances);
		}
		if (cudaGetLastError() != cudaSuccess) { kb_user_err = "kb_user_synth_process: launch failed"; return -2; }
		if (!dev && (cudaMemcpyAsync(out, d_result, floats * 4, cudaMemcpyDeviceToHost, stream) != cudaSuccess || cudaStreamSynchronize(stream) != cudaSuccess)) { kb_user_err = "kb_user_synth_process: D2H failed"; return -2; }
		return 0;
	}
	~kb_user_synth_bank() { cudaSetDevice(device); if (stream) cudaStreamSynchronize(stream); cudaFree(d_syn); cudaFree(d_notes); cudaFree(d_ctl); cudaFree(d_active); cudaFree(d_streams); cudaFree(d_out); if (stream) cudaStreamDestroy(stream); }
};

#define KB_USER_EXPORT_SYNTH(SYN, NOTE, NAME)                                                                                                \
	typedef kb_user_synth_bank<SYN, NOTE> kb_user_sbank;                                                                                     \
	extern "C" const char* kb_user_name(void) { return NAME; }                                                                               \
	extern "C" int kb_user_kind(void) { return 1; }                                                                                          \
	extern "C" int kb_user_channels(void) { return 1; }                                                                                      \
	extern "C" const char* kb_user_last_error(void) { return kb_user_err.c_str(); }                                                           \
	extern "C" int kb_user_num_controls(void) { std::vector<SYN> one(1); return one[0].controls.size(); }                                      \
	extern "C" int kb_user_synth_voices(void) { std::vector<SYN> one(1); return one[0].notes.count; }                                          \
	extern "C" void* kb_user_synth_create(int instances, float fs, int max_block, int device) {                                               \
		int ndev = 0;                                                                                                                         \
		if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) { cudaGetLastError(); kb_user_err = "kb_user_synth_create: no such CUDA device (there is no CPU path)"; return nullptr; } \
		if (instances < 1 || instances > 4096 || max_block < 1 || !(fs > 0.f)) { kb_user_err = "kb_user_synth_create: bad argument"; return nullptr; } \
		kb_user_sbank* b = new kb_user_sbank();                                                                                               \
		b->instances = instances; b->max_block = max_block; b->device = device; b->fs = kb_make_fs(fs);                                        \
		kb_kd_fs_host = b->fs;                                                                                                                \
		b->syn.resize(instances);                                                                                                             \
		b->voices = b->syn[0].notes.count;                                                                                                    \
		if (b->voices < 1 || b->voices > 128) { kb_user_err = "kb_user_synth_create: the program adds no notes (or more than 128)"; delete b; return nullptr; } \
		klang::kb_noise_registry.clear();                                                                                                     \
		b->notes.resize((size_t)instances * b->voices);                                                                                       \
		b->find_noise();                                                                                                                      \
		b->noteOns.assign(instances, 0u); b->noteStart.assign((size_t)instances * b->voices, 0u); b->dirty_note.assign((size_t)instances * b->voices, 1); \
		bool ok = cudaSetDevice(device) == cudaSuccess && cudaStreamCreateWithFlags(&b->stream, cudaStreamNonBlocking) == cudaSuccess &&       \
		          cudaMemcpyToSymbol(kb_kd_fs_dev, &b->fs, sizeof(KbFs)) == cudaSuccess &&                                                     \
		          cudaMalloc(&b->d_syn, sizeof(SYN) * instances) == cudaSuccess && cudaMalloc(&b->d_notes, sizeof(NOTE) * b->notes.size()) == cudaSuccess && \
		          cudaMalloc(&b->d_ctl, sizeof(klang::Controls) * instances) == cudaSuccess && cudaMalloc(&b->d_active, sizeof(int) * b->notes.size()) == cudaSuccess && \
		          cudaMalloc(&b->d_streams, sizeof(float) * b->notes.size() * max_block) == cudaSuccess &&                                     \
		          cudaMalloc(&b->d_out, sizeof(float) * (size_t)instances * max_block) == cudaSuccess &&                                       \
		          cudaMemcpy(b->d_syn, b->syn.data(), sizeof(SYN) * instances, cudaMemcpyHostToDevice) == cudaSuccess;                         \
		if (!ok) { kb_user_err = "kb_user_synth_create: CUDA allocation failed"; delete b; return nullptr; }                                   \
		return b;                                                                                                                             \
	}                                                                                                                                         \
	extern "C" void kb_user_synth_destroy(void* p) { delete static_cast<kb_user_sbank*>(p); }                                                  \
	extern "C" int kb_user_synth_set_control(void* p, int inst, int idx, float v) {                                                           \
		kb_user_sbank* b = static_cast<kb_user_sbank*>(p);                                                                                    \
		if (!b || inst < 0 || inst >= b->instances || idx < 0 || idx >= b->syn[0].controls.size()) { kb_user_err = "kb_user_synth_set_control: bad argument"; return -1; } \
		b->syn[inst].controls[idx].set(v); b->ctl_dirty = true; b->syn_dirty = true;                                                          \
		return 0;                                                                                                                             \
	}                                                                                                                                         \
	extern "C" int kb_user_synth_get_control(void* p, int inst, int idx, float* v) {                                                          \
		kb_user_sbank* b = static_cast<kb_user_sbank*>(p);                                                                                    \
		if (!b || !v || inst < 0 || inst >= b->instances || idx < 0 || idx >= b->syn[0].controls.size()) { kb_user_err = "kb_user_synth_get_control: bad argument"; return -1; } \
		*v = b->syn[inst].controls[idx].value;                                                                                                \
		return 0;                                                                                                                             \
	}                                                                                                                                         \
	extern "C" int kb_user_synth_voice_start(void* p, int inst, int voice, float pitch, float velocity) {                                     \
		kb_user_sbank* b = static_cast<kb_user_sbank*>(p);                                                                                    \
		if (!b || inst < 0 || inst >= b->instances || voice < 0 || voice >= b->voices) { kb_user_err = "kb_user_synth_voice_start: bad argument"; return -1; } \
		return b->start(inst, voice, pitch, velocity);                                                                                        \
	}                                                                                                                                         \
	extern "C" int kb_user_synth_voice_release(void* p, int inst, int voice, float velocity) {                                                \
		kb_user_sbank* b = static_cast<kb_user_sbank*>(p);                                                                                    \
		if (!b || inst < 0 || inst >= b->instances || voice < 0 || voice >= b->voices) { kb_user_err = "kb_user_synth_voice_release: bad argument"; return -1; } \
		return b->release(inst, voice, velocity);                                                                                             \
	}                                                                                                                                         \
	extern "C" int kb_user_synth_voice_stage(void* p, int inst, int voice) {                                                                  \
		kb_user_sbank* b = static_cast<kb_user_sbank*>(p);                                                                                    \
		if (!b || inst < 0 || inst >= b->instances || voice < 0 || voice >= b->voices) { kb_user_err = "kb_user_synth_voice_stage: bad argument"; return -1; } \
		if (b->fetch()) return -2;                                                                                                            \
		return b->note(inst, voice).stage;                                                                                                    \
	}                                                                                                                                         \
	/* Synth::noteOn / noteOff   klang.h:4423-4434 */                                                                                          \
	extern "C" int kb_user_synth_note_on(void* p, int inst, int pitch, float velocity) {                                                      \
		kb_user_sbank* b = static_cast<kb_user_sbank*>(p);                                                                                    \
		if (!b || inst < 0 || inst >= b->instances) { kb_user_err = "kb_user_synth_note_on: bad argument"; return -1; }                        \
		if (b->fetch()) return -2;                                                                                                            \
		const int v = b->assign(inst);                                                                                                        \
		const int rc = b->start(inst, v, (float)pitch, velocity);                                                                             \
		return rc ? rc : v;                                                                                                                   \
	}                                                                                                                                         \
	extern "C" int kb_user_synth_note_off(void* p, int inst, int pitch, float velocity) {                                                     \
		kb_user_sbank* b = static_cast<kb_user_sbank*>(p);                                                                                    \
		if (!b || inst < 0 || inst >= b->instances) { kb_user_err = "kb_user_synth_note_off: bad argument"; return -1; }                       \
		if (b->fetch()) return -2;                                                                                                            \
		for (int v = 0; v < b->voices; v++)                                                                                                   \
			if (b->note(inst, v).pitch == pitch && b->note(inst, v).stage == klang::Note::Sustain) { const int rc = b->release(inst, v, velocity); if (rc) return rc; } \
		return 0;                                                                                                                             \
	}                                                                                                                                         \
	extern "C" int kb_user_synth_process(void* p, float* out, int n, unsigned flags) {                                                        \
		kb_user_sbank* b = static_cast<kb_user_sbank*>(p);                                                                                    \
		if (!b) { kb_user_err = "kb_user_synth_process: null bank"; return -1; }                                                               \
		return b->process(out, n, flags);                                                                                                     \
	}

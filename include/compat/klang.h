// klang-b200 — a from-scratch `klang.h` for hosts that run klang programs on the B200.
//
// Purpose: an existing, UNMODIFIED `.k` program (`#include <klang.h>` … `struct X : Effect / Synth`) compiles against
// this header, and its block driver — Effect::process(buffer) / Synth::process(...) of the reference
// (nashaudio/klang klang.h:4203-4217, 4376-4467, 4703-4859) — is served by the hand-written sm_100a kernels of
// libklang_b200.so through the C ABI (include/klang_b200.h).  What the host needs from the `.k` object is what its
// constructor declares: the controls table (names, ranges, initial values), the presets and the number of notes; the
// per-sample `process()` bodies are type-checked here but evaluated on the device by the kernel of the same graph, and
// `Note::on()/off()` run inside the library where they draw libc rand() and call the host libm like the reference.
// The graph a plugin type maps to is declared once by the host with KLANG_B200_EFFECT / KLANG_B200_SYNTH (SURVEY H5
// "tier A": graphs selected by type).
//
// This is NOT the reference header and shares no code with it: value types are thin float wrappers whose operators exist so
// that the DSL expressions of the BASELINE programs are well-formed C++.  Every operator also has an honest, simple
// host meaning (signals hold floats, `a >> b` stores, controls clamp and smooth), but no block loop is implemented on
// the host — there is no CPU path.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <initializer_list>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../klang_b200.h"

namespace klang {

struct Control;

// constants with the reference's member spellings (`root2.f`, `root2.inv`, `x > root2`, `lfo.set(rate, pi)`)
struct constant {
	double d; float f; float inv;
	constexpr constant(double v) : d(v), f((float)v), inv((float)(1.0 / v)) {}
	constexpr operator float() const { return f; }
};
static constexpr constant pi(3.14159265358979323846), root2(1.41421356237309504880), ln2(0.69314718055994530942);

// ------------------------------------------------------------------------------------------------ values
struct signal {
	float value;
	signal(float v = 0.f) : value(v) {}
	signal(const constant& c) : value(c.f) {}
	signal(double v) : value((float)v) {}
	signal(int v) : value((float)v) {}
	signal(const Control& c);
	operator float() const { return value; }
	signal& operator=(float v) { value = v; return *this; }
	signal& operator+=(float v) { value += v; return *this; }
	signal& operator-=(float v) { value -= v; return *this; }
	signal& operator*=(float v) { value *= v; return *this; }
	signal& operator/=(float v) { value /= v; return *this; }
	signal& operator<<(float v) { value = v; return *this; }
};
struct param : signal {
	using signal::signal;
	param(const signal& s) : signal(s.value) {}
	param(const Control& c);
};
inline signal& operator>>(float x, signal& dst) { dst.value = x; return dst; }

struct SampleRate {
	float f; int i; double d; float inv, w, nyquist;
	SampleRate(float sr = 44100.f) : f(sr), i((int)(sr + 0.001f)), d(sr), inv(1.f / sr), w(2.f * pi.f / sr), nyquist(sr / 2.f) {}
	operator float() const { return f; }
};
static SampleRate fs;

template <class T> inline T sqr(T x) { return x * x; }
inline float sqr(const Control& c);
template <class T> inline T cube(T x) { return x * x * x; }
inline float cube(const Control& c);
inline float random(float lo, float hi) { return rand() * ((hi - lo) / (float)RAND_MAX) + lo; }
inline double random(double lo, double hi) { return rand() * ((hi - lo) / (double)RAND_MAX) + lo; }
inline void random(unsigned seed) { srand(seed); }

// Pitch -> Frequency
struct Conversion { param Frequency; };
struct Pitch : signal {
	using signal::signal;
	const Conversion* operator->() const { static thread_local Conversion c; c.Frequency = kb_pitch_to_frequency(value); return &c; }
};
typedef signal Amplitude;
typedef signal Velocity;
typedef void event;

struct Debug { signal last; };
static Debug debug;
inline Debug& operator>>(float x, Debug& d) { d.last = x; return d; }

// ---------------------------------------------------------------------------------------------- controls
struct Control {
	std::string name;
	float min = 0.f, max = 1.f, initial = 0.f, value = 0.f;
	signal smoothed;
	std::vector<std::string> options;
	struct Size { int x, y, w, h; } size = { 0, 0, 0, 0 };
	Control() {}
	operator float() const { return value; }
	void set(float x) { value = x < min ? min : (max < x ? max : x); }
	float smooth() { smoothed = smoothed * 0.999f + (1.f - 0.999f) * value; return smoothed; }
protected:
	Control(const char* n, float lo, float hi, float init) : name(n), min(lo), max(hi), initial(init), value(init) {}
};
inline signal::signal(const Control& c) : value(c.value) {}
inline param::param(const Control& c) : signal(c.value) {}
inline float sqr(const Control& c) { return c.value * c.value; }
inline float cube(const Control& c) { return c.value * c.value * c.value; }

struct Dial : Control { Dial(const char* n, float lo = 0.f, float hi = 1.f, float init = 0.f) : Control(n, lo, hi, init) {} };
struct Slider : Control { Slider(const char* n, float lo = 0.f, float hi = 1.f, float init = 0.f) : Control(n, lo, hi, init) {} };
struct Toggle : Control { Toggle(const char* n, float init = 0.f) : Control(n, 0.f, 1.f, init) {} };
struct Button : Control { Button(const char* n) : Control(n, 0.f, 1.f, 0.f) {} };
struct Meter : Control { Meter(const char* n, float lo = 0.f, float hi = 1.f, float init = 0.f) : Control(n, lo, hi, init) {} };
struct Menu : Control {
	template <class... Opts> Menu(const char* n, Opts... opts) : Control(n, 0.f, (float)(sizeof...(Opts)) - 1.f, 0.f) { options = { std::string(opts)... }; }
};
struct Group {
	std::string name; std::vector<Control> items;
	template <class... C> Group(const char* n, const C&... c) : name(n) { items = { static_cast<const Control&>(c)... }; }
};
struct Controls {
	std::vector<Control> items;
	Controls& operator=(std::initializer_list<Control> list) { items.assign(list.begin(), list.end()); return *this; }
	Controls& operator=(std::initializer_list<Group> groups) { items.clear(); for (const Group& g : groups) items.insert(items.end(), g.items.begin(), g.items.end()); return *this; }
	Control& operator[](int i) { if ((size_t)i >= items.size()) items.resize(i + 1); return items[i]; }
	int size() const { return (int)items.size(); }
	bool changed() { bool c = false; cached.resize(items.size(), 0.f); for (size_t i = 0; i < items.size(); i++) if (items[i].value != cached[i]) { cached[i] = items[i].value; c = true; } return c; }
private:
	std::vector<float> cached;
};
struct Preset { std::string name; std::vector<float> values; Preset(const char* n, std::initializer_list<float> v) : name(n), values(v) {} };
struct Presets {
	std::vector<Preset> items;
	Presets& operator=(std::initializer_list<Preset> list) { items.assign(list.begin(), list.end()); return *this; }
};

// ------------------------------------------------------------------------------------ dataflow protocol
// an object that yields a sample when read (conversion == one tick), and one that also accepts a sample
struct Generator {
	signal out;
	virtual ~Generator() {}
	virtual void process() {}
	operator float() { process(); return out; }
	operator signal() { process(); return out; }
	signal operator++(int) { process(); return out; }
};
struct Modifier : Generator {
	signal in;
	virtual void input() {}
	template <class... P> Modifier& operator()(P... p) { set(param(p)...); return *this; }   // `in >> lpf(f, Q) >> out` (klang.h:2306-2308)
	virtual void set(param) {}
	virtual void set(param, param) {}
	virtual void set(param, param, param) {}
	virtual void set(param, param, param, param) {}
};
inline Modifier& operator>>(float x, Modifier& m) { m.in = x; m.input(); return m; }

struct Oscillator : Generator {
	param frequency = 1000.f, phase = 0.f, duty = 0.5f;
	void set(param f) { frequency = f; }
	void set(param f, param p) { frequency = f; phase = p; }
	void set(param f, param p, param d) { frequency = f; phase = p; duty = d; }
	template <class... P> Oscillator& operator()(P... p) { set(param(p)...); return *this; }   // `lfo(rate) * depth`: set, then read (klang.h:2258-2260)
	void reset() { phase = 0.f; }
};

struct Envelope : Generator {
	struct Point { double x, y; Point(double x_ = 0, double y_ = 0) : x(x_), y(y_) {} };
	std::vector<Point> points;
	bool released = false;
	Envelope() {}
	Envelope(std::initializer_list<Point> p) : points(p) {}
	Envelope& operator=(std::initializer_list<Point> p) { points.assign(p.begin(), p.end()); return *this; }
	void release(float time = 0.f, float level = 0.f) { (void)time; (void)level; released = true; }
	bool finished() const { return released; }
	float at(param t) const { (void)t; return points.empty() ? 0.f : (float)points[0].y; }
	void setLoop(int, int) {}
};
struct ADSR : Envelope {
	param A = 0.5f, D = 0.5f, S = 1.f, R = 0.5f;
	void set(param a, param d, param s, param r) { A = a; D = d; S = s; R = r; }
	void operator()(param a, param d, param s, param r) { set(a, d, s, r); }
};

// FM operator (klang.h:4140-4173): an oscillator with an input, an envelope and an amplitude; `a * i >> b` feeds b's input
template <class OSCILLATOR> struct Operator : OSCILLATOR {
	signal in; Envelope env; Amplitude amp = 1.f;
	Operator& operator()(param f) { OSCILLATOR::set(f); return *this; }
	Operator& operator()(param f, param phase) { OSCILLATOR::set(f, phase); return *this; }
	Operator& operator=(std::initializer_list<Envelope::Point> p) { env = p; return *this; }
	Operator& operator=(const Envelope& e) { env = e; return *this; }
	Operator& operator*(signal a) { amp = a; return *this; }
	Operator& operator>>(Operator& carrier) { carrier.in = float(*this); return carrier; }
};

// lookup table filled from a function or a list (klang.h Table; FM.k:18-22), and the debug graph a Note may draw into
#define FUNCTION(type) (void(*)(type, type&))[](type x, type& y)
template <class T, int SIZE> struct Table {
	T v[SIZE];
	Table(void (*fn)(T, T&)) { for (int i = 0; i < SIZE; i++) { v[i] = T(); fn((T)i, v[i]); } }
	Table(T (*fn)(T)) { for (int i = 0; i < SIZE; i++) v[i] = fn((T)i); }
	Table(std::initializer_list<T> l) { int i = 0; for (const T& x : l) if (i < SIZE) v[i++] = x; for (; i < SIZE; i++) v[i] = T(); }
	T operator[](int i) const { return v[i]; }
};
struct Graph {                           // UI only: `f >> graph(x0, x1, y0, y1)` plots a function (klang.h:2536-2840); nothing is evaluated here
	void clear() {}
	template <class T> void add(const T&) {}
	Graph& operator()(double, double) { return *this; }
	Graph& operator()(double, double, double, double) { return *this; }
};
inline Graph& operator>>(float (*)(float), Graph& g) { return g; }
static Graph graph;
typedef param Frequency;

template <int SIZE> struct Delay : Modifier {
	param time = 1.f;
	using Modifier::set;
	void set(param samples) override { time = samples; }
	signal operator()(float delay) const { (void)delay; return signal(0.f); }
	Delay& operator<<(float x) { in = x; return *this; }
};
template <int SIZE> inline Delay<SIZE>& operator>>(float x, Delay<SIZE>& d) { d.in = x; return d; }

namespace Filters {
	struct Filter : Modifier {
		param f = 0.f, Q = 0.f;
		using Modifier::set;
		void set(param f_) override { f = f_; }
		void set(param f_, param Q_) override { f = f_; Q = Q_; }
		void reset() { f = 0.f; Q = 0.f; }
	};
	namespace Biquad { struct LPF : Filter {}; struct HPF : Filter {}; struct BPF : Filter {}; struct BRF : Filter {}; struct APF : Filter {}; }
	namespace OnePole { struct LPF : Filter {}; struct HPF : Filter {}; }
}
namespace OnePole = Filters::OnePole;

namespace Generators {
	namespace Fast { struct Sine : Oscillator {}; struct Saw : Oscillator {}; struct Triangle : Oscillator {}; struct Square : Oscillator {}; struct Pulse : Oscillator {}; struct Noise : Generator {}; }
	namespace Basic { struct Sine : Oscillator {}; struct Saw : Oscillator {}; struct Triangle : Oscillator {}; struct Square : Oscillator {}; struct Pulse : Oscillator {}; struct Noise : Generator {}; }
}

// fixed-capacity array with a live count (Reverb.k: `times.count = ...`)
template <class T, int CAPACITY> struct Array {
	T items[CAPACITY]; unsigned count = 0;
	T& operator[](int i) { return items[i]; }
	const T& operator[](int i) const { return items[i]; }
	void add(const T& x) { if (count < (unsigned)CAPACITY) items[count++] = x; }
	unsigned size() const { return count; }
};

// N-channel sample and the 4x4 feedback matrix of Reverb.k's FDN
template <int N> struct signals {
	signal value[N];
	signals() {}
	template <class... A> signals(A&&... a) : value{ signal(static_cast<A&&>(a))... } {}
	signal& operator[](int i) { return value[i]; }
	const signal& operator[](int i) const { return value[i]; }
};
template <int N> inline signals<N> operator+(const signals<N>& a, float b) { signals<N> r; for (int i = 0; i < N; i++) r.value[i] = a.value[i] + b; return r; }
struct Matrix {
	float v[4][4];
	template <class... A> constexpr Matrix(A... a) : v{ { 0 } } { const float f[] = { (float)a... }; for (int i = 0; i < (int)sizeof...(A) && i < 16; i++) v[i / 4][i % 4] = f[i]; }
};
inline signals<4> operator>>(const signals<4>& in, const Matrix& m) {
	signals<4> out;
	for (int r = 0; r < 4; r++) { float acc = 0.f; for (int c = 0; c < 4; c++) acc += m.v[r][c] * in.value[c]; out.value[r] = acc; }
	return out;
}

// sample cursors over caller-owned memory (what Note::process(buffer) receives)
struct buffer {
	float* samples = nullptr; int size = 0; float* ptr = nullptr;
	buffer() {}
	buffer(float* data, int n) : samples(data), size(n), ptr(data) {}
	void rewind() { ptr = samples; }
	bool finished() const { return ptr >= samples + size; }
	signal& operator++(int) { return *reinterpret_cast<signal*>(ptr++); }
	buffer& operator+=(float x) { *ptr += x; return *this; }
	buffer& operator=(float x) { *ptr = x; return *this; }
	operator float() const { return *ptr; }
};

namespace Mono {
	typedef klang::signal signal;
	typedef klang::buffer buffer;
	typedef klang::Generator Generator;
	typedef klang::Modifier Modifier;
	typedef klang::Oscillator Oscillator;
}
namespace mono = Mono;

// ------------------------------------------------------------------------------------------ plugin shells
struct Plugin { Controls controls; Presets presets; virtual ~Plugin() {} };

struct Effect : Plugin {
	signal in, out;
	virtual void prepare() {}
	virtual void process() {}
};

struct NoteBase {
	Controls* synth_controls = nullptr;
	struct ControlsRef { NoteBase* n; Control& operator[](int i) { return (*n->synth_controls)[i]; } } controls{ this };
	bool stopped = false;
	virtual ~NoteBase() {}
	virtual void process() {}
	void stop() { stopped = true; }
	bool finished() const { return stopped; }
};
struct Note : NoteBase { signal out; };

struct Notes {
	Controls* owner = nullptr;
	std::vector<std::unique_ptr<NoteBase>> items;
	template <class NOTE> void add(int count) { for (int i = 0; i < count; i++) { items.emplace_back(new NOTE()); items.back()->synth_controls = owner; } }
	int size() const { return (int)items.size(); }
};
struct Synth : Effect {
	Notes notes;
	Synth() { notes.owner = &controls; }
};

// ------------------------------------------------------------------------------------------------ stereo
namespace Stereo {
	struct frame {
		klang::signal l, r;
		frame() {}
		frame(klang::signal l_, klang::signal r_) : l(l_), r(r_) {}
		frame& operator=(float v) { l = v; r = v; return *this; }
		frame& operator+=(const frame& x) { l += x.l; r += x.r; return *this; }
		frame& operator<<(const frame& x) { l = x.l; r = x.r; return *this; }
	};
	typedef frame signal;
	inline frame operator*(const frame& a, const frame& b) { return frame(a.l * b.l, a.r * b.r); }
	inline frame operator*(const frame& a, float b) { return frame(a.l * b, a.r * b); }
	inline frame operator*(float a, const frame& b) { return frame(a * b.l, a * b.r); }
	inline frame operator+(const frame& a, const frame& b) { return frame(a.l + b.l, a.r + b.r); }
	inline frame& operator>>(const frame& a, frame& dst) { dst = a; return dst; }

	struct Generator {
		frame out;
		virtual ~Generator() {}
		virtual void process() {}
		operator frame() { process(); return out; }
	};
	struct Modifier : Generator {
		frame in;
		virtual void set(param) {}
		virtual void set(param, param) {}
		virtual void set(param, param, param) {}
		virtual void set(param, param, param, param) {}
	};
	inline Modifier& operator>>(const frame& x, Modifier& m) { m.in = x; return m; }
	inline frame operator*(Modifier& m, float g) { return (frame)m * g; }
	struct Oscillator : Generator {
		param frequency = 1000.f;
		virtual void set(param) {}
		virtual void set(param, param) {}
		virtual void set(param, param, param) {}
	};
	// N parallel copies of a mono modifier, one per channel (Reverb.k: Stereo::Bank<LPF>)
	template <class T> struct Bank : Modifier {
		T items[2];
		void set(param f) override { items[0].set(f); items[1].set(f); }
		void set(param f, param q) override { items[0].set(f, q); items[1].set(f, q); }
	};
	struct buffer {
		klang::buffer left, right;
		buffer() {}
		buffer(float* l, float* r, int n) : left(l, n), right(r, n) {}
		void rewind() { left.rewind(); right.rewind(); }
		bool finished() const { return left.finished(); }
		klang::buffer& channel(int c) { return c ? right : left; }
		// `buffer++ *= env` (SynTHX.k:178) scales the frame the cursor just left, in place
		struct Cursor {
			float* l; float* r;
			Cursor& operator*=(float g) { *l *= g; *r *= g; return *this; }
			Cursor& operator=(const frame& x) { *l = x.l; *r = x.r; return *this; }
			Cursor& operator+=(const frame& x) { *l += x.l; *r += x.r; return *this; }
			operator frame() const { return frame(*l, *r); }
		};
		Cursor operator++(int) { Cursor c{ left.ptr, right.ptr }; left.ptr++; right.ptr++; return c; }
		buffer& operator+=(const frame& x) { *left.ptr += x.l; *right.ptr += x.r; return *this; }
	};

	template <int SIZE> struct Delay {
		klang::Delay<SIZE> l, r;
		frame operator()(const frame& delay) const { (void)delay; return frame(); }
		frame operator()(float delay) const { (void)delay; return frame(); }
		Delay& operator<<(const frame& x) { l.in = x.l; r.in = x.r; return *this; }
	};
	template <int SIZE> inline Delay<SIZE>& operator>>(const frame& x, Delay<SIZE>& d) { return d << x; }
	template <int SIZE> inline Delay<SIZE>& operator>>(Modifier& m, Delay<SIZE>& d) { return d << (frame)m; }
	inline Modifier& operator>>(Modifier& a, Modifier& b) { b.in = (frame)a; return b; }

	struct Effect : klang::Plugin {
		frame in, out;
		virtual void prepare() {}
		virtual void process() {}
	};
	struct Note : klang::NoteBase { frame out; virtual bool process(buffer) { return !finished(); } using klang::NoteBase::process; };
	struct Synth : Effect {
		klang::Notes notes;
		Synth() { notes.owner = &controls; }
	};
}
namespace stereo = Stereo;

namespace optimised { using namespace Generators::Fast; using namespace Filters::Biquad; }
namespace basic { using namespace Generators::Basic; using namespace Filters::Biquad; }
namespace minimal { using namespace Generators::Basic; using namespace Filters::Biquad; }

// =============================================================================================== B200 hosts
namespace b200 {

template <class PLUGIN> struct graph_of;      // specialised by KLANG_B200_EFFECT / KLANG_B200_SYNTH

// The device runs a hand-written restatement of ONE reference program per graph id; this header only type-checks a `.k` body, it does not
// evaluate it.  So a program is bound to an id only when the hash of the `.k` source that was actually compiled (tools/k_hash.py, passed
// by the build: tools/build_k_host.py) equals the hash of the program the id restates (kb_graph_source_hash): an edited body, or another
// plugin with the same controls table, is refused instead of silently running the old DSP.
inline void check_source(bool synth, int graph, unsigned long long compiled_hash) {
	const unsigned long long want = kb_graph_source_hash(synth ? 1 : 0, graph);
	if (want == 0ULL || compiled_hash != want)
		throw std::logic_error(std::string("the .k program compiled here is not the program graph ") + std::to_string(graph) + " restates (" + kb_graph_source_path(synth ? 1 : 0, graph) +
		                       "): source hash mismatch — the device would run different DSP than the program describes; refusing to bind");
}

struct Error : std::runtime_error { explicit Error(const std::string& what) : std::runtime_error(what + ": " + kb_last_error()) {} };

// Owns one PLUGIN object (for its controls / presets, as the reference host does) and the bank that evaluates it.
// process() is Effect::process(buffer) / Stereo::Effect::process: control values are pushed like
// `controls[c].set(params[c])`, the block is processed in place on the device, values the graph modified are read back
// (klang.h:4208-4216, 4444-4447, 4462-4465).
template <class PLUGIN> class EffectHost {
public:
	PLUGIN plugin;
	EffectHost(float sample_rate, int max_block, int instances = 1, int device = 0)
		: bank_((check_source(false, graph_of<PLUGIN>::id, graph_of<PLUGIN>::source_hash), kb_fx_bank_create(graph_of<PLUGIN>::id, instances, sample_rate, max_block, device))) {
		if (!bank_) throw Error("kb_fx_bank_create");
		const int n = kb_fx_bank_num_controls(bank_);
		if (n != plugin.controls.size()) throw std::logic_error("controls table of the .k program does not match the bound graph");
		for (int c = 0; c < n; c++) {
			float v = 0.f;
			kb_fx_bank_get_control(bank_, 0, c, &v);
			if (v != plugin.controls[c].value) throw std::logic_error("control " + plugin.controls[c].name + ": initial value differs from the bound graph");
		}
		pushed_.assign(n, 0.f);
		for (int c = 0; c < n; c++) pushed_[c] = plugin.controls[c].value;
	}
	~EffectHost() { kb_fx_bank_destroy(bank_); }
	EffectHost(const EffectHost&) = delete;
	int channels() const { return kb_fx_bank_channels(bank_); }
	// buffer: planar [instances][channels][length], in place
	bool process(float* buffer, int length) {
		for (int c = 0; c < (int)pushed_.size(); c++)
			if (plugin.controls[c].value != pushed_[c]) {
				for (int i = 0; i < kb_fx_bank_instances(bank_); i++) kb_fx_bank_set_control(bank_, i, c, plugin.controls[c].value);
				kb_fx_bank_get_control(bank_, 0, c, &pushed_[c]);
				plugin.controls[c].value = pushed_[c];
			}
		return kb_fx_bank_process(bank_, buffer, length, 0) == KB_OK;
	}
	// params[c] = controls[c].value after the block (only PingPong.k writes a control per sample)
	void readControls() { for (int c = 0; c < (int)pushed_.size(); c++) { kb_fx_bank_get_control(bank_, 0, c, &pushed_[c]); plugin.controls[c].value = pushed_[c]; } }
	kb_fx_bank* bank() { return bank_; }
private:
	kb_fx_bank* bank_;
	std::vector<float> pushed_;
};

// Synth::process(float*, int) / Stereo::Synth::process plus noteOn / noteOff (klang.h:4423-4466, 4830-4858).
template <class PLUGIN> class SynthHost {
public:
	PLUGIN plugin;
	SynthHost(float sample_rate, int max_block, int instances = 1, int device = 0)
		: bank_((check_source(true, graph_of<PLUGIN>::id, graph_of<PLUGIN>::source_hash), kb_synth_bank_create(graph_of<PLUGIN>::id, instances, plugin.notes.size() > 0 ? plugin.notes.size() : 32, sample_rate, max_block, device))) {
		if (!bank_) throw Error("kb_synth_bank_create");
		const int n = kb_synth_bank_num_controls(bank_);
		if (n != plugin.controls.size()) throw std::logic_error("controls table of the .k program does not match the bound graph");
		for (int c = 0; c < n; c++) {
			float v = 0.f;
			kb_synth_bank_get_control(bank_, 0, c, &v);
			if (v != plugin.controls[c].value) throw std::logic_error("control " + plugin.controls[c].name + ": initial value differs from the bound graph");
		}
		pushed_.assign(n, 0.f);
		for (int c = 0; c < n; c++) pushed_[c] = plugin.controls[c].value;
	}
	~SynthHost() { kb_synth_bank_destroy(bank_); }
	SynthHost(const SynthHost&) = delete;
	int channels() const { return kb_synth_bank_channels(bank_); }
	int voices() const { return kb_synth_bank_voices(bank_); }
	int noteOn(int pitch, float velocity, int instance = 0) { sync(); return kb_synth_bank_note_on(bank_, instance, pitch, velocity); }
	void noteOff(int pitch, float velocity = 0.f, int instance = 0) { sync(); kb_synth_bank_note_off(bank_, instance, pitch, velocity); }
	// Synth::input(status, byte1, byte2): raw MIDI (templates/juce/synth/Source/klang.h:3921-3929)
	void input(int status, int byte1, int byte2, int instance = 0) { sync(); kb_synth_bank_midi(bank_, instance, status, byte1, byte2); }
	bool process(float* buffer, int length, unsigned flags = 0) { sync(); return kb_synth_bank_process(bank_, buffer, length, flags) == KB_OK; }
	kb_synth_bank* bank() { return bank_; }
private:
	void sync() {
		for (int c = 0; c < (int)pushed_.size(); c++)
			if (plugin.controls[c].value != pushed_[c]) {
				for (int i = 0; i < kb_synth_bank_instances(bank_); i++) kb_synth_bank_set_control(bank_, i, c, plugin.controls[c].value);
				kb_synth_bank_get_control(bank_, 0, c, &pushed_[c]);
				plugin.controls[c].value = pushed_[c];
			}
	}
	kb_synth_bank* bank_;
	std::vector<float> pushed_;
};

}  // namespace b200
}  // namespace klang

// SOURCE_HASH = the hash of the `.k` file PLUGIN was compiled from (python tools/k_hash.py file.k prints the #define)
#define KLANG_B200_EFFECT(PLUGIN, GRAPH, SOURCE_HASH) namespace klang { namespace b200 { template <> struct graph_of<PLUGIN> { static constexpr int id = GRAPH; static constexpr unsigned long long source_hash = SOURCE_HASH; }; } }
#define KLANG_B200_SYNTH(PLUGIN, GRAPH, SOURCE_HASH) namespace klang { namespace b200 { template <> struct graph_of<PLUGIN> { static constexpr int id = GRAPH; static constexpr unsigned long long source_hash = SOURCE_HASH; }; } }

using namespace klang;

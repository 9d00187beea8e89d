// klang-b200 — C ABI (include/klang_b200.h) over the kernels of kb_kernels.cuh.
//
// Host side of a bank: a mirror of the POD state for the event-rate code (Note::on/off, Notes::assign,
// Control::set, Effect::prepare), kept coherent with the device copy by two flags:
//   host_stale  the device ran a block since the mirror was last fetched  -> fetch before the next event
//   dirty       the mirror was modified by events                         -> upload before the next block
// Blocks never touch the host mirror and never run on the host.
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <utility>
#include <vector>

#include "../../include/klang_b200.h"
#include "kb_kernels.cuh"
#include "kb_tiled.cuh"
#include "kb_reverb3.cuh"
#include "kb_pingpong3.cuh"
#include "kb_k_hashes.h"
#include "kb_presets.h"

static thread_local std::string g_err = "";
static int kb_fail(int code, const std::string& msg) { g_err = msg; return code; }
#define KB_CUDA(call)                                                                              \
	do {                                                                                           \
		cudaError_t e_ = (call);                                                                   \
		if (e_ != cudaSuccess) return kb_fail(KB_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
	} while (0)

extern "C" int kb_version(void) { return KB_VERSION; }
extern "C" const char* kb_last_error(void) { return g_err.c_str(); }
extern "C" int kb_device_count(void) { int n = 0; if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; } return n; }
extern "C" void kb_srand(unsigned seed) { srand(seed); }
extern "C" float kb_pitch_to_frequency(float pitch) { return kb_pitch_to_frequency_host(pitch); }
// which klang program a graph id restates (kb_k_hashes.h): a host binds its compiled `.k` to an id only when the source hashes agree
extern "C" unsigned long long kb_graph_source_hash(int synth, int graph) {
	if (graph < 0 || graph >= (synth ? KB_SY_COUNT : KB_FX_COUNT)) return 0ULL;
	return (synth ? kb_k_synth_sources : kb_k_fx_sources)[graph].hash;
}
extern "C" const char* kb_graph_source_path(int synth, int graph) {
	if (graph < 0 || graph >= (synth ? KB_SY_COUNT : KB_FX_COUNT)) return "";
	return (synth ? kb_k_synth_sources : kb_k_fx_sources)[graph].path;
}

template <class T> static cudaError_t dev_alloc(T** p, size_t count) { return cudaMalloc((void**)p, count * sizeof(T) > 0 ? count * sizeof(T) : 1); }

struct kb_bank_base {
	int device = 0;
	cudaStream_t stream = nullptr, own_stream = nullptr;
	long long launches = 0;
	KbFs fs;
	int max_block = 0;
	float* d_io = nullptr; size_t io_floats = 0;   // staging for host-pointer calls
	float* d_old = nullptr;                        // Flanger.k / Chorus.k: what the block's write sweep overwrote, [instances][max_block]
	bool host_stale = false, dirty = true;
	// measurement: CUDA events around the dominant kernel of each process() call
	bool profiling = false; std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof_events; size_t prof_used = 0;
	void prof_begin() {
		if (!profiling) return;
		if (prof_used == prof_events.size()) { cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b); prof_events.push_back({ a, b }); }
		cudaEventRecord(prof_events[prof_used].first, stream);
	}
	void prof_end() { if (profiling) cudaEventRecord(prof_events[prof_used++].second, stream); }
	int prof_read(double* ms, long long* count) {
		if (cudaStreamSynchronize(stream) != cudaSuccess) return KB_ECUDA;
		double sum = 0;
		for (size_t i = 0; i < prof_used; i++) { float t = 0; cudaEventElapsedTime(&t, prof_events[i].first, prof_events[i].second); sum += t; }
		if (ms) *ms = sum;
		if (count) *count = (long long)prof_used;
		return KB_OK;
	}
	void prof_free() { for (auto& e : prof_events) { cudaEventDestroy(e.first); cudaEventDestroy(e.second); } prof_events.clear(); prof_used = 0; }
};

// multi-GPU mix-down state (the entry points are further down: "multi-GPU mix-down")
struct kb_mixdown {
	int device = 0, world = 1, rank = 0, max_floats = 0;
	unsigned step = 0;                  // steps acquired so far on this rank
	unsigned collected = 0;             // rank 0: the last step whose slots have been summed
	unsigned char* arena = nullptr;     // rank 0: own allocation; others: the IPC mapping
	unsigned* d_tickets = nullptr;      // local device memory: two CTA tickets of the fused step kernel (publish, collect)
	cudaEvent_t ev_done = nullptr;      // recorded behind the last fused step kernel (which may run on a bank's side stream)
	cudaEvent_t ev_ring[4] = { nullptr, nullptr, nullptr, nullptr };   // the same per step (step & 3): kb_mixdown_host_wait joins an older step
	bool step_pending = false;
	bool mapped = false;
	size_t slot_bytes() const { return (size_t)max_floats * sizeof(float); }
	float* slot(unsigned step_, int r) const { return (float*)(arena + ((size_t)(step_ & 1u) * world + r) * slot_bytes()); }
	volatile unsigned* flags() const { return (volatile unsigned*)(arena + 2 * (size_t)world * slot_bytes()); }
	volatile unsigned* consumed() const { return flags() + world; }
	size_t arena_bytes() const { return 2 * (size_t)world * slot_bytes() + sizeof(unsigned) * ((size_t)world + 1); }
};

// =============================================================================================== effects
struct kb_fx_bank : kb_bank_base {
	int graph = 0, instances = 0, channels = 0, ncontrols = 0;
	size_t state_bytes = 0; long long ring_floats = 0;
	std::vector<KbFxHdr> hdr; std::vector<unsigned char> state;
	KbFxHdr* d_hdr = nullptr; unsigned char* d_state = nullptr; float* d_rings = nullptr; KbFxPlan* d_plan = nullptr; void* d_sync = nullptr; int epoch = 0; std::vector<KbFxPlan> plan_cache;
	bool device_writes_controls = false;
	unsigned last_flags = 0; int last_schedule = 0;             // flags / Reverb.k schedule of the last process() call
	bool rv_all_resident = false;                               // Reverb.k: every instance runs on kb_reverb3_kernel (no fallback launches needed)
	float* d_debug = nullptr; bool debug_on = false; int debug_n = -1;   // `>> debug` capture of the last block, [instances][max_block] (kb_fx_bank_debug_*)
	template <class T> T& st(int i) { return *reinterpret_cast<T*>(state.data() + (size_t)i * state_bytes); }
};

static int fx_fetch(kb_fx_bank* b) {
	if (!b->host_stale) return KB_OK;
	KB_CUDA(cudaSetDevice(b->device));
	KB_CUDA(cudaMemcpyAsync(b->hdr.data(), b->d_hdr, b->hdr.size() * sizeof(KbFxHdr), cudaMemcpyDeviceToHost, b->stream));
	KB_CUDA(cudaMemcpyAsync(b->state.data(), b->d_state, b->state.size(), cudaMemcpyDeviceToHost, b->stream));
	KB_CUDA(cudaStreamSynchronize(b->stream));
	b->host_stale = false;
	return KB_OK;
}
static int fx_upload(kb_fx_bank* b) {
	if (!b->dirty) return KB_OK;
	KB_CUDA(cudaMemcpyAsync(b->d_hdr, b->hdr.data(), b->hdr.size() * sizeof(KbFxHdr), cudaMemcpyHostToDevice, b->stream));
	KB_CUDA(cudaMemcpyAsync(b->d_state, b->state.data(), b->state.size(), cudaMemcpyHostToDevice, b->stream));
	// the mirror may be modified again before the copy engine has read it
	KB_CUDA(cudaStreamSynchronize(b->stream));
	b->dirty = false;
	return KB_OK;
}

extern "C" kb_fx_bank* kb_fx_bank_create(int graph, int instances, float fs, int max_block, int device) {
	if (graph < 0 || graph >= KB_FX_COUNT || instances < 1 || max_block < 1 || !(fs > 0)) { kb_fail(KB_EINVAL, "kb_fx_bank_create: bad argument"); return nullptr; }
	// (several kernels carry the row index instance * channels + channel in gridDim.y, which is limited to 65535)
	if (instances > 32767) { kb_fail(KB_EINVAL, "kb_fx_bank_create: at most 32767 instances per bank (create several banks)"); return nullptr; }
	if (device < 0 || device >= kb_device_count()) { kb_fail(KB_ENODEV, "kb_fx_bank_create: no such CUDA device (klang-b200 has no CPU path)"); return nullptr; }
	kb_fx_bank* b = new kb_fx_bank();
	b->graph = graph; b->instances = instances; b->device = device; b->max_block = max_block; b->fs = kb_make_fs(fs);
	switch (graph) {
	case KB_FX_GAIN: b->channels = 1; b->ncontrols = 1; b->state_bytes = sizeof(KbGainFx); b->ring_floats = 0; break;
	case KB_FX_PINGPONG: b->channels = 2; b->ncontrols = 6; b->state_bytes = sizeof(KbPingPong); b->ring_floats = KB_PINGPONG_RING_FLOATS; b->device_writes_controls = true; break;
	case KB_FX_REVERB: b->channels = 2; b->ncontrols = 10; b->state_bytes = sizeof(KbReverb); b->ring_floats = KB_REVERB_RING_FLOATS; break;
	case KB_FX_DELAY_PINGPONG: b->channels = 2; b->ncontrols = 4; b->state_bytes = sizeof(KbDPingPong); b->ring_floats = KB_PINGPONG_RING_FLOATS; break;
	case KB_FX_DELAY_REVERB: b->channels = 1; b->ncontrols = 3; b->state_bytes = sizeof(KbDReverb); b->ring_floats = KB_PINGPONG_RING_FLOATS; break;
	case KB_FX_PAN: b->channels = 2; b->ncontrols = 1; b->state_bytes = sizeof(KbGainFx); b->ring_floats = 0; break;
	case KB_FX_RM: case KB_FX_TREMOLO: b->channels = 1; b->ncontrols = 2; b->state_bytes = sizeof(KbLfoFx); b->ring_floats = 0; break;
	case KB_FX_CLIPPING: case KB_FX_FUNCTIONS: case KB_FX_MUTE: b->channels = 1; b->ncontrols = 1; b->state_bytes = sizeof(KbGainFx); b->ring_floats = 0; break;
	case KB_FX_ECHO: case KB_FX_FEEDBACK: b->channels = 1; b->ncontrols = 2; b->state_bytes = sizeof(KbOneDelayFx); b->ring_floats = KB_ONEDELAY_RING_FLOATS; break;
	case KB_FX_IIR: b->channels = 1; b->ncontrols = 1; b->state_bytes = sizeof(KbIirFx); b->ring_floats = 0; break;
	case KB_FX_WAHWAH: b->channels = 1; b->ncontrols = 3; b->state_bytes = sizeof(KbWahWahFx); b->ring_floats = 0; break;
	case KB_FX_FLANGER: case KB_FX_MODDELAY: case KB_FX_MOD_CHORUS: b->channels = 1; b->ncontrols = 2; b->state_bytes = sizeof(KbModDelayFx); b->ring_floats = KB_ONEDELAY_RING_FLOATS; break;
	}
	b->hdr.assign(instances, KbFxHdr());
	memset(b->hdr.data(), 0, b->hdr.size() * sizeof(KbFxHdr));
	b->state.assign((size_t)instances * b->state_bytes, 0);
	for (int i = 0; i < instances; i++) {
		const long long ring0 = (long long)i * b->ring_floats;
		switch (graph) {
		case KB_FX_GAIN: b->hdr[i].controls[0] = kb_dial(0.f, 1.f, 0.5f); break;                  // Gain.k:10
		case KB_FX_PINGPONG: kb_pingpong_construct(b->hdr[i], b->st<KbPingPong>(i), ring0); break;
		case KB_FX_REVERB: kb_reverb_construct(b->hdr[i], b->st<KbReverb>(i), ring0); break;
		case KB_FX_DELAY_PINGPONG: kb_dpingpong_construct(b->hdr[i], b->st<KbDPingPong>(i), ring0); break;
		case KB_FX_DELAY_REVERB: kb_dreverb_construct(b->hdr[i], b->st<KbDReverb>(i), ring0); break;
		case KB_FX_PAN: b->hdr[i].controls[0] = kb_dial(0.f, 1.f, 0.5f); break;                                               // Pan.k:10
		case KB_FX_RM: case KB_FX_TREMOLO:                                                                                    // RM.k:11-12, Tremolo.k:11-12
			b->hdr[i].controls[0] = kb_dial(1.f, graph == KB_FX_RM ? 1000.f : 10.f, 6.f); b->hdr[i].controls[1] = kb_dial(0.f, 0.5f, 0.5f);
			kb_fsine_init(b->st<KbLfoFx>(i).lfo); break;
		case KB_FX_CLIPPING: b->hdr[i].controls[0] = kb_dial(1.f, 11.f, 1.f); break;                                          // Clipping.k:10
		case KB_FX_FUNCTIONS: b->hdr[i].controls[0] = kb_dial(1.f, 25.f, 1.f); break;                                         // Functions.k:18
		case KB_FX_MUTE: b->hdr[i].controls[0] = kb_dial(0.f, 1.f, 0.f); break;                                               // Toggle("Mute"), Mute.k:10
		case KB_FX_IIR: b->hdr[i].controls[0] = kb_dial(0.f, 1.f, 0.5f); b->st<KbIirFx>(i).last = 0.f; break;                  // IIR.k:5, 10
		case KB_FX_FLANGER: case KB_FX_MODDELAY: case KB_FX_MOD_CHORUS: {                                                     // Flanger.k:11-14, ModDelay.k:11-14, Chorus.k:11-14
			KbModDelayFx& m = b->st<KbModDelayFx>(i);
			kb_delay_construct(m.delay, 192000, ring0);
			for (int k = 0; k < 3; k++) kb_fsine_init(m.lfo[k]);
			kb_osm_construct(m.tri, 0, 1.0f);                                                                                 // Fast::Triangle (saw waveform, duty 1)
			if (graph == KB_FX_FLANGER) { b->hdr[i].controls[0] = kb_dial(0.1f, 1.0f, 0.75f); b->hdr[i].controls[1] = kb_dial(0.1f, 5.0f, 1.5f); }
			else if (graph == KB_FX_MODDELAY) { b->hdr[i].controls[0] = kb_dial(1.f, 10.f, 6.f); b->hdr[i].controls[1] = kb_dial(0.f, 1.f, 0.2f); }
			else { b->hdr[i].controls[0] = kb_dial(1.f, 10.f, 6.f); b->hdr[i].controls[1] = kb_dial(0.f, 1.f, 0.1f); }
			break; }
		case KB_FX_WAHWAH:                                                                                                    // WahWah.k:10-14
			b->hdr[i].controls[0] = kb_dial(10.f, 10000.f, 1000.f); b->hdr[i].controls[1] = kb_dial(0.1f, 10.f, 1.f); b->hdr[i].controls[2] = kb_dial(4.f, 10.f, 6.f);
			kb_biquad_construct(b->st<KbWahWahFx>(i).lpf, KB_BQ_LPF); kb_fsine_init(b->st<KbWahWahFx>(i).lfo); break;
		case KB_FX_ECHO: case KB_FX_FEEDBACK:                                                                                 // Echo.k:10-13, Feedback.k:10-13
			b->hdr[i].controls[0] = kb_dial(0.f, 1.f, 0.5f); b->hdr[i].controls[1] = kb_dial(0.f, 1.f, 0.5f);
			kb_delay_construct(b->st<KbOneDelayFx>(i).delay, 192000, ring0); break;
		}
	}
	bool ok = cudaSetDevice(device) == cudaSuccess;
	ok = ok && cudaStreamCreateWithFlags(&b->own_stream, cudaStreamNonBlocking) == cudaSuccess;
	b->stream = b->own_stream;
	ok = ok && dev_alloc(&b->d_hdr, instances) == cudaSuccess;
	ok = ok && dev_alloc(&b->d_state, b->state.size()) == cudaSuccess;
	ok = ok && dev_alloc(&b->d_rings, (size_t)instances * b->ring_floats) == cudaSuccess;
	ok = ok && dev_alloc(&b->d_plan, instances) == cudaSuccess;
	ok = ok && cudaMalloc(&b->d_sync, sizeof(KbDppSync) + sizeof(int) * (size_t)instances * KB_DPP_MAXCHUNKS) == cudaSuccess;
	ok = ok && cudaMemsetAsync(b->d_sync, 0, sizeof(KbDppSync) + sizeof(int) * (size_t)instances * KB_DPP_MAXCHUNKS, b->stream) == cudaSuccess;
	ok = ok && cudaMemsetAsync(b->d_plan, 0, instances * sizeof(KbFxPlan), b->stream) == cudaSuccess;
	ok = ok && cudaFuncSetAttribute(kb_reverb_par_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(KbRvSmem)) == cudaSuccess;
	ok = ok && cudaFuncSetAttribute(kb_reverb_pipe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(KbRv2Smem)) == cudaSuccess;
	ok = ok && cudaFuncSetAttribute(kb_reverb3_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(KbRv3Smem)) == cudaSuccess;
	ok = ok && cudaFuncSetAttribute(kb_reverb3_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(KbRv3Smem)) == cudaSuccess;
	if (ok && b->ring_floats) ok = cudaMemsetAsync(b->d_rings, 0, (size_t)instances * b->ring_floats * sizeof(float), b->stream) == cudaSuccess;
	b->io_floats = (size_t)instances * b->channels * max_block;
	ok = ok && dev_alloc(&b->d_io, b->io_floats) == cudaSuccess;
	if (graph == KB_FX_FLANGER || graph == KB_FX_MOD_CHORUS) ok = ok && dev_alloc(&b->d_old, (size_t)instances * max_block) == cudaSuccess;
	if (graph == KB_FX_MODDELAY) ok = ok && dev_alloc(&b->d_old, 2 * (size_t)instances * max_block) == cudaSuccess;      // stash + the per-frame depth rows
	if (!ok) { kb_fail(KB_ECUDA, std::string("kb_fx_bank_create: ") + cudaGetErrorString(cudaGetLastError())); kb_fx_bank_destroy(b); return nullptr; }
	return b;
}
extern "C" void kb_fx_bank_destroy(kb_fx_bank* b) {
	if (!b) return;
	cudaSetDevice(b->device);
	if (b->stream) cudaStreamSynchronize(b->stream);
	cudaFree(b->d_hdr); cudaFree(b->d_state); cudaFree(b->d_rings); cudaFree(b->d_io); cudaFree(b->d_old); cudaFree(b->d_plan); cudaFree(b->d_sync); cudaFree(b->d_debug);
	b->prof_free();
	if (b->own_stream) cudaStreamDestroy(b->own_stream);
	delete b;
}
extern "C" int kb_fx_bank_channels(const kb_fx_bank* b) { return b ? b->channels : KB_EINVAL; }
extern "C" int kb_fx_bank_instances(const kb_fx_bank* b) { return b ? b->instances : KB_EINVAL; }
extern "C" int kb_fx_bank_num_controls(const kb_fx_bank* b) { return b ? b->ncontrols : KB_EINVAL; }
extern "C" long long kb_fx_bank_launches(const kb_fx_bank* b) { return b ? b->launches : 0; }
extern "C" int kb_fx_bank_parallel_instances(kb_fx_bank* b) {
	if (!b) return kb_fail(KB_EINVAL, "null bank");
	if (b->graph == KB_FX_GAIN || (b->graph >= KB_FX_PAN && b->graph <= KB_FX_CLIPPING) || b->graph == KB_FX_FUNCTIONS || b->graph == KB_FX_MUTE) return b->instances;
	if (b->graph == KB_FX_IIR || b->graph == KB_FX_WAHWAH) return 0;                 // a recurrence: one lane per instance
	if (b->graph >= KB_FX_FLANGER && b->graph <= KB_FX_MOD_CHORUS) return b->instances;
	if (b->graph == KB_FX_ECHO) return b->instances;                                 // (blocks longer than SIZE - fs frames fall back to the sequential schedule)
	if (b->graph == KB_FX_FEEDBACK) {                                                // instances whose delay is long enough for a chunk (at this block size)
		int count = 0;
		for (int i = 0; i < b->instances; i++) count += kb_feedback_chunk(b->fs, b->max_block, b->hdr[i].controls[0].value) > 0;
		return count;
	}
	std::vector<KbFxPlan> plan(b->instances);
	KB_CUDA(cudaSetDevice(b->device));
	KB_CUDA(cudaStreamSynchronize(b->stream));
	KB_CUDA(cudaMemcpy(plan.data(), b->d_plan, plan.size() * sizeof(KbFxPlan), cudaMemcpyDeviceToHost));
	int count = 0;
	for (const KbFxPlan& p : plan) count += (p.mode & KB_PLAN_PARALLEL) != 0;
	return count;
}
extern "C" int kb_fx_bank_tolerance_instances(kb_fx_bank* b) {
	if (!b) return kb_fail(KB_EINVAL, "null bank");
	if (b->graph != KB_FX_REVERB || !(b->last_flags & KB_FX_TOLERANCE) || (b->last_flags & KB_FX_SEQUENTIAL)) return 0;
	std::vector<KbFxPlan> plan(b->instances);
	KB_CUDA(cudaSetDevice(b->device));
	KB_CUDA(cudaStreamSynchronize(b->stream));
	KB_CUDA(cudaMemcpy(plan.data(), b->d_plan, plan.size() * sizeof(KbFxPlan), cudaMemcpyDeviceToHost));
	int count = 0;
	for (const KbFxPlan& p : plan) count += (p.mode & KB_PLAN_PARALLEL) && (p.mode & KB_PLAN_SCAN_OK);
	return b->last_schedule == 3 ? count : 0;
}
extern "C" long long kb_fx_bank_state_bytes(const kb_fx_bank* b) { return b ? (long long)(b->hdr.size() * sizeof(KbFxHdr) + b->state.size()) : 0; }
extern "C" int kb_fx_bank_profile(kb_fx_bank* b, int enable) { if (!b) return kb_fail(KB_EINVAL, "null bank"); cudaSetDevice(b->device); cudaStreamSynchronize(b->stream); b->profiling = enable != 0; b->prof_used = 0; return KB_OK; }
extern "C" int kb_fx_bank_profile_read(kb_fx_bank* b, double* ms, long long* count) { if (!b) return kb_fail(KB_EINVAL, "null bank"); cudaSetDevice(b->device); return b->prof_read(ms, count); }
extern "C" int kb_fx_bank_sync(kb_fx_bank* b) { if (!b) return kb_fail(KB_EINVAL, "null bank"); KB_CUDA(cudaSetDevice(b->device)); KB_CUDA(cudaStreamSynchronize(b->stream)); return KB_OK; }
extern "C" int kb_fx_bank_set_stream(kb_fx_bank* b, void* s) {
	if (!b) return kb_fail(KB_EINVAL, "null bank");
	KB_CUDA(cudaStreamSynchronize(b->stream));
	b->stream = s ? (cudaStream_t)s : b->own_stream;
	return KB_OK;
}
extern "C" int kb_fx_bank_set_control(kb_fx_bank* b, int inst, int idx, float v) {
	if (!b || inst < 0 || inst >= b->instances || idx < 0 || idx >= b->ncontrols) return kb_fail(KB_EINVAL, "kb_fx_bank_set_control: bad argument");
	// the upload replaces the whole blob, so the mirror must be current before it is modified
	int rc = fx_fetch(b); if (rc) return rc;
	kb_control_set(b->hdr[inst].controls[idx], v);
	b->dirty = true;
	return KB_OK;
}
extern "C" int kb_fx_bank_get_control(kb_fx_bank* b, int inst, int idx, float* v) {
	if (!b || !v || inst < 0 || inst >= b->instances || idx < 0 || idx >= b->ncontrols) return kb_fail(KB_EINVAL, "kb_fx_bank_get_control: bad argument");
	if (b->device_writes_controls) { int rc = fx_fetch(b); if (rc) return rc; }
	*v = b->hdr[inst].controls[idx].value;
	return KB_OK;
}
// Factory presets (kb_presets.h) -------------------------------------------------------------------------------------------------
extern "C" int kb_graph_num_presets(int is_synth, int graph) { return kb_preset_count(is_synth ? 1 : 0, graph); }
extern "C" int kb_graph_preset(int is_synth, int graph, int index, char* name, int name_max, float* values, int max_values) {
	const KbPreset* p = kb_preset_find(is_synth ? 1 : 0, graph, index);
	if (!p || index < 0) return kb_fail(KB_EINVAL, "kb_graph_preset: no such preset");
	if (name && name_max > 0) { strncpy(name, p->name, (size_t)name_max - 1); name[name_max - 1] = 0; }
	for (int c = 0; values && c < p->count && c < max_values; c++) values[c] = p->values[c];
	return p->count;
}
// the host's preset load: every value through Control::set (klang.h:1725-1728), then onPreset (klang.h:4190 — no bound program overrides preset())
extern "C" int kb_fx_bank_load_preset(kb_fx_bank* b, int inst, int index) {
	if (!b || inst < 0 || inst >= b->instances) return kb_fail(KB_EINVAL, "kb_fx_bank_load_preset: bad argument");
	const KbPreset* p = index >= 0 ? kb_preset_find(0, b->graph, index) : nullptr;
	if (!p) return kb_fail(KB_EINVAL, "kb_fx_bank_load_preset: the bank's program has no such preset");
	for (int c = 0; c < p->count && c < b->ncontrols; c++) { int rc = kb_fx_bank_set_control(b, inst, c, p->values[c]); if (rc) return rc; }
	return KB_OK;
}
extern "C" double kb_fx_bank_bytes_per_frame(kb_fx_bank* b) {
	if (!b) return 0;
	switch (b->graph) {
	case KB_FX_GAIN: return 8;
	case KB_FX_PINGPONG: return 48;                                // 16 io + 2 x (4 write + 12 read), SURVEY §8d
	case KB_FX_REVERB: { fx_fetch(b); int taps = 10 + (int)(b->hdr[0].controls[6].value * (float)10.999); return 16 + 2 * (4 + taps * 8) + 16 * 20; }
	case KB_FX_DELAY_PINGPONG: return 40;
	case KB_FX_DELAY_REVERB: return 88;
	case KB_FX_PAN: return 16;
	case KB_FX_RM: case KB_FX_TREMOLO: case KB_FX_CLIPPING: case KB_FX_FUNCTIONS: case KB_FX_MUTE: return 8;
	case KB_FX_IIR: case KB_FX_WAHWAH: return 8;
	case KB_FX_FLANGER: case KB_FX_MODDELAY: return 20;
	case KB_FX_MOD_CHORUS: return 36;                               // 8 io + 4 write + 3 taps x 8
	case KB_FX_ECHO: case KB_FX_FEEDBACK: return 20;               // 8 io + 4 write + two adjacent floats read
	}
	return 0;
}

static int fx_prepare(kb_fx_bank* b) {
	// Effect::prepare() of every instance (klang.h:4209), on the host mirror
	if (b->graph == KB_FX_PINGPONG) {
		// dcfilter.set(50,1) is idempotent: it only does work the first time (PingPong.k:36-40, klang.h:5588)
		bool need = false;
		for (int i = 0; i < b->instances && !need; i++) { const KbPingPong& p = b->st<KbPingPong>(i); need = p.dc[0].f != 50.f || p.dc[0].Q != 1.f; }
		if (need) {
			int rc = fx_fetch(b); if (rc) return rc;
			for (int i = 0; i < b->instances; i++) kb_pingpong_prepare(b->fs, b->st<KbPingPong>(i));
			b->dirty = true;
		}
	} else if (b->graph == KB_FX_REVERB) {
		bool changed = false;
		for (int i = 0; i < b->instances && !changed; i++)
			for (int c = 0; c < 10; c++) if (b->hdr[i].controls[c].value != b->hdr[i].cached[c]) changed = true;
		if (changed) {
			int rc = fx_fetch(b); if (rc) return rc;
			for (int i = 0; i < b->instances; i++) kb_reverb_prepare(b->fs, b->hdr[i], b->st<KbReverb>(i));
			b->dirty = true;
			// the schedule of an instance depends only on what prepare() sets (read-to-write distances, damping coefficients, tap times): evaluate
			// the kernel's own plan on the mirror, so the fallback kernels are launched only for banks that need them
			b->rv_all_resident = true;
			for (int i = 0; i < b->instances; i++) {
				const KbReverb& rv = b->st<KbReverb>(i);
				KbRv3LinePlan lines[16];
				for (int line = 0; line < 16; line++) lines[line] = kb_rv3_plan_line((line < 8 ? rv.mid[line >> 2] : rv.late[(line - 8) >> 2]).d[line & 3]);
				const KbFxPlan p = kb_rv3_plan_combine(lines, rv.times, rv.count, rv.dl.SIZE, rv.dr.SIZE);
				if (!(p.mode & KB_PLAN_PARALLEL) || !(p.mode & KB_PLAN_RESIDENT)) b->rv_all_resident = false;
			}
		}
	} else if (b->graph == KB_FX_RM || b->graph == KB_FX_TREMOLO) {
		// `lfo(rate)` = Fast::Sine::set(rate), a no-op while the rate equals the cached frequency (RM.k:21, Tremolo.k:26, klang.h:5143-5147, Q3).
		// The device only ever advances the phase, so the mirror's frequency is current without a fetch.
		bool need = false;
		for (int i = 0; i < b->instances && !need; i++) need = b->st<KbLfoFx>(i).lfo.frequency != b->hdr[i].controls[0].value;
		if (need) {
			int rc = fx_fetch(b); if (rc) return rc;
			for (int i = 0; i < b->instances; i++) kb_fsine_set_f(b->fs, b->st<KbLfoFx>(i).lfo, b->hdr[i].controls[0].value);
			b->dirty = true;
		}
	} else if (b->graph == KB_FX_DELAY_REVERB) {
		// filter.set(controls[2]) (Delay/Reverb.k:59-61)
		bool need = false;
		for (int i = 0; i < b->instances && !need; i++) { const KbDReverb& p = b->st<KbDReverb>(i); need = p.filter.f != b->hdr[i].controls[2].value || p.filter.Q != KB_ROOT2_INV_F; }
		if (need) {
			int rc = fx_fetch(b); if (rc) return rc;
			for (int i = 0; i < b->instances; i++) kb_biquad_set_f(b->fs, b->st<KbDReverb>(i).filter, b->hdr[i].controls[2].value);
			b->dirty = true;
		}
	}
	return KB_OK;
}

// Debug taps (include/klang_b200.h).  While enabled, every process() also leaves the block's `>> debug` capture on the device.
extern "C" int kb_fx_bank_debug_enable(kb_fx_bank* b, int enable) {
	if (!b) return kb_fail(KB_EINVAL, "kb_fx_bank_debug_enable: null bank");
	KB_CUDA(cudaSetDevice(b->device));
	if (enable && !b->d_debug) KB_CUDA(cudaMalloc(&b->d_debug, (size_t)b->instances * b->max_block * sizeof(float)));
	b->debug_on = enable != 0; b->debug_n = -1;
	return KB_OK;
}
extern "C" int kb_fx_bank_debug_read(kb_fx_bank* b, float* dst, int n, unsigned flags) {
	if (!b || !dst || n < 0) return kb_fail(KB_EINVAL, "kb_fx_bank_debug_read: bad argument");
	if (!b->debug_on) return kb_fail(KB_EINVAL, "kb_fx_bank_debug_read: debug capture is not enabled (kb_fx_bank_debug_enable)");
	if (b->debug_n < 0) return 0;                              // the program has no tap, or no block since the last read (Buffer::get, klang.h:3164-3172)
	if (n != b->debug_n) return kb_fail(KB_EINVAL, "kb_fx_bank_debug_read: n differs from the last block's length");
	KB_CUDA(cudaSetDevice(b->device));
	KB_CUDA(cudaMemcpy2DAsync(dst, (size_t)n * sizeof(float), b->d_debug, (size_t)b->max_block * sizeof(float), (size_t)n * sizeof(float), b->instances,
	                          (flags & KB_DEVICE_PTR) ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, b->stream));
	if (!(flags & KB_DEVICE_PTR)) KB_CUDA(cudaStreamSynchronize(b->stream));
	b->debug_n = -1;
	return 1;
}

// CTAs per row of a streaming kernel (256 threads, 4 x 16 bytes in flight per thread): one trip per thread — many short CTAs that the hardware
// scheduler balances, like a library copy.  A grid of long-running CTAs that is a few CTAs larger than one wave leaves those few running alone,
// latency-bound, behind the wave.  KB_STREAM_WAVE=1 (A/B measurement): at most one wave of 148 SMs x 8 resident CTAs, grid-stride loops.
static unsigned kb_stream_grid_x(int n, int rows, int vectors_per_trip = 4) {
	static const bool one_wave = getenv("KB_STREAM_WAVE") && atoi(getenv("KB_STREAM_WAVE")) != 0;
	const int cover = (n / (4 * vectors_per_trip) + 255) / 256;
	const int wave = (148 * 8) / std::max(1, std::min(rows, 148 * 8));
	return (unsigned)std::max(1, one_wave ? std::min(cover, wave) : cover);
}
extern "C" int kb_fx_bank_process(kb_fx_bank* b, float* io, int n, unsigned flags) {
	if (!b || !io || n < 0 || n > b->max_block) return kb_fail(KB_EINVAL, "kb_fx_bank_process: bad argument (n > max_block?)");
	if (n == 0) return KB_OK;
	KB_CUDA(cudaSetDevice(b->device));
	int rc = fx_prepare(b); if (rc) return rc;
	rc = fx_upload(b); if (rc) return rc;
	const size_t floats = (size_t)b->instances * b->channels * n;
	float* d = io;
	if (!(flags & KB_DEVICE_PTR)) { d = b->d_io; KB_CUDA(cudaMemcpyAsync(d, io, floats * sizeof(float), cudaMemcpyHostToDevice, b->stream)); }
	const int ib = (b->instances + 31) / 32;
	b->debug_n = -1;
	if (b->debug_on) {                                         // the block's `>> debug` capture, from the state the block starts with (kb_kernels.cuh)
		if (b->graph == KB_FX_PINGPONG || b->graph == KB_FX_MODDELAY) {
			kb_debug_tap_kernel<<<ib, 32, 0, b->stream>>>(b->graph, b->d_hdr, b->d_state, b->state_bytes, b->d_debug, n, b->max_block, b->instances, b->fs);
			b->debug_n = n; b->launches++;
		} else if (b->graph == KB_FX_RM || b->graph == KB_FX_TREMOLO) {
			dim3 grid((unsigned)std::max(1, std::min((n + 255) / 256, 64)), b->instances);
			kb_debug_tap_lfo_kernel<<<grid, 256, 0, b->stream>>>(b->graph, b->d_hdr, (const KbLfoFx*)b->d_state, b->d_debug, n, b->max_block);
			b->debug_n = n; b->launches++;
		}
	}
	b->prof_begin();
	b->last_flags = flags;
	const bool seq_only = flags & KB_FX_SEQUENTIAL;
	switch (b->graph) {
	case KB_FX_GAIN: {
		dim3 grid(kb_stream_grid_x(n, b->instances), b->instances);
		kb_gain_kernel<<<grid, 256, 0, b->stream>>>(b->d_hdr, d, n);
		break; }
	case KB_FX_PAN: case KB_FX_RM: case KB_FX_TREMOLO: case KB_FX_CLIPPING: case KB_FX_FUNCTIONS: case KB_FX_MUTE: {
		const int rows = b->instances * b->channels;
		const bool lfo = b->graph == KB_FX_RM || b->graph == KB_FX_TREMOLO;
		dim3 grid(kb_stream_grid_x(n, rows), rows);
		const KbLfoFx* lfos = lfo ? (const KbLfoFx*)b->d_state : nullptr;
		switch (b->graph) {
#define KB_EW(G) case G: kb_elementwise_kernel<G><<<grid, 256, 0, b->stream>>>(b->channels, b->d_hdr, lfos, d, n, n); break
		KB_EW(KB_FX_PAN); KB_EW(KB_FX_RM); KB_EW(KB_FX_TREMOLO); KB_EW(KB_FX_CLIPPING); KB_EW(KB_FX_FUNCTIONS); KB_EW(KB_FX_MUTE);
#undef KB_EW
		}
		if (lfo) { kb_lfo_advance_kernel<<<ib, 32, 0, b->stream>>>((KbLfoFx*)b->d_state, b->instances, n); b->launches++; }
		break; }
	case KB_FX_ECHO: {
		bool echo_par = !seq_only;                             // (controls are host-authoritative for this graph: the mirror is current)
		for (int i = 0; i < b->instances && echo_par; i++) echo_par = kb_echo_parallel_ok(b->fs, n, b->hdr[i].controls[0].value);
		if (echo_par) {                                        // time-parallel: write sweep, read sweep, position advance (kb_graphs.cuh)
			KbOneDelayFx* st = (KbOneDelayFx*)b->d_state;
			dim3 grid(kb_stream_grid_x(n, b->instances, 1), b->instances);      // (one frame per thread and trip, four trips: n / 4 threads per row)
			kb_echo_write_kernel<<<grid, 256, 0, b->stream>>>(st, b->d_rings, d, n, n);
			kb_echo_read_kernel<<<grid, 256, 0, b->stream>>>(b->d_hdr, st, b->d_rings, d, n, n, b->fs);
			kb_onedelay_advance_kernel<<<ib, 32, 0, b->stream>>>(st, b->instances, n);
			b->launches += 2;
		} else {                                               // one lane per instance, frame by frame
			kb_fx_seq_kernel<KB_FX_ECHO, KbOneDelayFx><<<ib, 32, 0, b->stream>>>(b->d_hdr, (KbOneDelayFx*)b->d_state, b->d_rings, d, n, n, 1, b->instances, b->fs, nullptr);
		}
		break; }
	case KB_FX_IIR:        // one lane per instance (the smoother is a serial fp32 chain)
		kb_fx_seq_kernel<KB_FX_IIR, KbIirFx><<<ib, 32, 0, b->stream>>>(b->d_hdr, (KbIirFx*)b->d_state, b->d_rings, d, n, n, 1, b->instances, b->fs, nullptr);
		break;
	case KB_FX_FLANGER: case KB_FX_MOD_CHORUS: case KB_FX_MODDELAY:
		if (!seq_only && n < 192000) {                         // time-parallel: write sweep with stash, read sweep (kb_modline_*, kb_graphs.cuh)
			KbModDelayFx* st = (KbModDelayFx*)b->d_state;
			dim3 grid(kb_stream_grid_x(n, b->instances, 1), b->instances);
			float* depth = b->graph == KB_FX_MODDELAY ? b->d_old + (size_t)b->instances * b->max_block : nullptr;   // ModDelay.k: serial smoother pre-pass
			kb_modline_begin_kernel<<<ib, 32, 0, b->stream>>>(b->graph, b->d_hdr, st, b->instances, b->fs, depth, n, n);
			kb_modline_write_kernel<<<grid, 256, 0, b->stream>>>(st, b->d_rings, b->d_old, d, n, n);
			grid.x = kb_stream_grid_x(2 * n, b->instances, 1);                  // (two trips per thread)
			kb_modline_read_kernel<<<grid, 256, 0, b->stream>>>(b->graph, b->d_hdr, st, b->d_rings, b->d_old, depth, d, n, n, b->fs);
			kb_modline_end_kernel<<<ib, 32, 0, b->stream>>>(b->graph, st, b->instances, n);
			b->launches += 3;
		} else if (b->graph == KB_FX_MODDELAY) {               // one lane per instance, frame by frame
			kb_fx_seq_kernel<KB_FX_MODDELAY, KbModDelayFx><<<ib, 32, 0, b->stream>>>(b->d_hdr, (KbModDelayFx*)b->d_state, b->d_rings, d, n, n, 1, b->instances, b->fs, nullptr);
		} else if (b->graph == KB_FX_FLANGER) {
			kb_fx_seq_kernel<KB_FX_FLANGER, KbModDelayFx><<<ib, 32, 0, b->stream>>>(b->d_hdr, (KbModDelayFx*)b->d_state, b->d_rings, d, n, n, 1, b->instances, b->fs, nullptr);
		} else {
			kb_fx_seq_kernel<KB_FX_MOD_CHORUS, KbModDelayFx><<<ib, 32, 0, b->stream>>>(b->d_hdr, (KbModDelayFx*)b->d_state, b->d_rings, d, n, n, 1, b->instances, b->fs, nullptr);
		}
		break;
	case KB_FX_WAHWAH:     // one lane per instance (the biquad state is a serial fp32 chain; the coefficients could come from parallel workers as in C2)
		kb_fx_seq_kernel<KB_FX_WAHWAH, KbWahWahFx><<<ib, 32, 0, b->stream>>>(b->d_hdr, (KbWahWahFx*)b->d_state, b->d_rings, d, n, n, 1, b->instances, b->fs, nullptr);
		break;
	case KB_FX_FEEDBACK:   // chunk-parallel (chunks shorter than the delay, one CTA per instance) unless KB_FX_SEQUENTIAL
		if (!seq_only) kb_feedback_par_kernel<<<b->instances, 1024, 0, b->stream>>>(b->d_hdr, (KbOneDelayFx*)b->d_state, b->d_rings, d, n, n, b->fs);
		else kb_fx_seq_kernel<KB_FX_FEEDBACK, KbOneDelayFx><<<ib, 32, 0, b->stream>>>(b->d_hdr, (KbOneDelayFx*)b->d_state, b->d_rings, d, n, n, 1, b->instances, b->fs, nullptr);
		break;
	case KB_FX_PINGPONG: {
		KbPingPong* st = (KbPingPong*)b->d_state;
		// sub-blocks of at most 8192 frames (the staged block lives in shared memory); every sub-block is planned on the device
		// KB_PP_SCHEDULE (A/B measurement; same results): 3 = default, ONE fused launch per sub-block (kb_pingpong3.cuh: in-kernel plan, producers eight
		// chunks ahead of the filter lanes, sequential fallback inside); 2 = the round-1 plan / network / filter / finish launches
		static const int pp_schedule = getenv("KB_PP_SCHEDULE") ? atoi(getenv("KB_PP_SCHEDULE")) : 3;
		for (int o = 0; o < n && pp_schedule == 3 && !seq_only; o += 8192) {
			const int len = std::min(8192, n - o);
			kb_pingpong3_kernel<<<b->instances, KB_PP3_NT, 0, b->stream>>>(b->d_hdr, st, b->d_plan, b->d_rings, d + o, len, n, b->fs);
			if (o + 8192 < n) b->launches++;
		}
		for (int o = 0; o < n && (pp_schedule != 3 || seq_only); o += 8192) {
			const int len = std::min(8192, n - o);
			if (!seq_only) {
				kb_pingpong_plan_kernel<<<ib, 32, 0, b->stream>>>(b->d_hdr, st, b->d_plan, b->instances, len, b->fs);
				kb_pingpong_par_kernel<1024><<<b->instances * 2, 1024, (((size_t)len + 7) / 8 * 8 + 8) * sizeof(float), b->stream>>>(b->d_hdr, st, b->d_plan, b->d_rings, d + o, len, n, b->fs);
				kb_pingpong_finish_kernel<<<ib, 32, 0, b->stream>>>(st, b->d_plan, b->instances, len);
				b->launches += 3;
			}
			kb_fx_seq_kernel<KB_FX_PINGPONG, KbPingPong><<<ib, 32, 0, b->stream>>>(b->d_hdr, st, b->d_rings, d + o, len, n, 2, b->instances, b->fs, seq_only ? nullptr : b->d_plan);
			if (o + 8192 < n) b->launches++;
		}
		break; }
	case KB_FX_REVERB: {
		KbReverb* st = (KbReverb*)b->d_state;
		// KB_RV_SCHEDULE (A/B measurement; same results): 3 = default, decoupled roles around bulk-async staging (kb_reverb3.cuh);
		// 2 = the round-1 software pipeline (one CTA barrier per chunk); 1 = the unpipelined chunk kernel
		static const int rv_env = getenv("KB_RV_SCHEDULE") ? atoi(getenv("KB_RV_SCHEDULE")) : 3;
		// the bulk copies of schedule 3 need 16-byte aligned io rows
		const int rv_schedule = (rv_env == 3 && ((reinterpret_cast<uintptr_t>(d) & 15) != 0 || (n & 3) != 0)) ? 2 : rv_env;
		b->last_schedule = rv_schedule;
		const int sub = 1 << 20;                       // sub-blocks keep the kernels' tick counters in 32 bits
		for (int o = 0; o < n; o += sub) {
			const int len = std::min(sub, n - o);
			if (!seq_only) {
				if (rv_schedule == 3) {
					// (the plan of every instance is made inside the kernel by its own CTAs and left in d_plan for the kernels launched behind it)
					const int tol = (flags & KB_FX_TOLERANCE) ? 1 : 0;
					// KB_FX_TOLERANCE: instances whose line filters the scan admits run on the tolerance kernel, the others on the exact one
					// KB_RV3_TRACE=<file> (measurement aid): per-role, per-chunk clock64() stamps of CTA 0 of every launch, last launch wins
					static const char* trace_path = getenv("KB_RV3_TRACE");
					static long long* d_trace = nullptr;
					const size_t trace_n = (size_t)KB_RV3_TRACE_ROLES * KB_RV3_TRACE_CHUNKS * 2;
					if (trace_path && !d_trace) { KB_CUDA(cudaMalloc(&d_trace, trace_n * sizeof(long long))); }
					if (d_trace) KB_CUDA(cudaMemsetAsync(d_trace, 0, trace_n * sizeof(long long), b->stream));
					if (tol) { kb_reverb3_kernel<1><<<b->instances * 2, KB_RV3_NT_TOL, sizeof(KbRv3Smem), b->stream>>>(b->d_hdr, st, b->d_plan, b->d_rings, d + o, len, n, 0, d_trace); b->launches++; }
					kb_reverb3_kernel<0><<<b->instances * 2, KB_RV3_NT, sizeof(KbRv3Smem), b->stream>>>(b->d_hdr, st, b->d_plan, b->d_rings, d + o, len, n, tol, tol ? nullptr : d_trace);
					if (d_trace) {
						std::vector<long long> tr(trace_n);
						KB_CUDA(cudaMemcpyAsync(tr.data(), d_trace, trace_n * sizeof(long long), cudaMemcpyDeviceToHost, b->stream));
						KB_CUDA(cudaStreamSynchronize(b->stream));
						if (FILE* f = fopen(trace_path, "w")) {
							for (int r = 0; r < KB_RV3_TRACE_ROLES; r++) for (int k = 0; k < KB_RV3_TRACE_CHUNKS; k++)
								if (tr[((size_t)r * KB_RV3_TRACE_CHUNKS + k) * 2]) fprintf(f, "%d %d %lld %lld\n", r, k, tr[((size_t)r * KB_RV3_TRACE_CHUNKS + k) * 2], tr[((size_t)r * KB_RV3_TRACE_CHUNKS + k) * 2 + 1]);
							fclose(f);
						}
					}
					// instances whose live delay spans do not fit in shared memory (fs = 192 kHz) keep the round-1 pipeline (its CTAs exit at once otherwise)
					if (!b->rv_all_resident) {
						kb_reverb_pipe_kernel<<<b->instances * 2, KB_RV2_NT, sizeof(KbRv2Smem), b->stream>>>(b->d_hdr, st, b->d_plan, b->d_rings, d + o, len, n);
						b->launches++;
					}
				} else if (rv_schedule == 1) {
					kb_reverb_plan_kernel<<<ib, 32, 0, b->stream>>>(st, b->d_plan, b->instances);
					kb_reverb_par_kernel<<<b->instances * 2, 256, sizeof(KbRvSmem), b->stream>>>(b->d_hdr, st, b->d_plan, b->d_rings, d + o, len, n);
				} else {
					kb_reverb_plan2_kernel<<<ib, 32, 0, b->stream>>>(st, b->d_plan, b->instances);
					kb_reverb_pipe_kernel<<<b->instances * 2, KB_RV2_NT, sizeof(KbRv2Smem), b->stream>>>(b->d_hdr, st, b->d_plan, b->d_rings, d + o, len, n);
				}
				b->launches += 2;
			}
			if (seq_only || rv_schedule != 3 || !b->rv_all_resident) {
				kb_fx_seq_kernel<KB_FX_REVERB, KbReverb><<<ib, 32, 0, b->stream>>>(b->d_hdr, st, b->d_rings, d + o, len, n, 2, b->instances, b->fs, seq_only ? nullptr : b->d_plan);
				if (o + sub < n) b->launches++;
			} else b->launches--;                         // (the common accounting below counts one launch for the sequential kernel)
		}
		break; }
	case KB_FX_DELAY_PINGPONG: {
		KbDPingPong* st = (KbDPingPong*)b->d_state;
		// the plan of this graph depends on host-owned controls only: made on the host, uploaded when it changes
		static const int cf = getenv("KB_DPP_CHUNK") ? atoi(getenv("KB_DPP_CHUNK")) : 1024;
		bool all_parallel = !seq_only;
		if (!seq_only) {
			std::vector<KbFxPlan> plan(b->instances);
			for (int i = 0; i < b->instances; i++) {
				const float tl = b->hdr[i].controls[0].value * b->fs.f, tr = b->hdr[i].controls[1].value * b->fs.f;
				KbFxPlan& p = plan[i];
				p.chunk = (int)std::min(tl, tr) - 2;
				// far-end guard: a frame of this launch must never overwrite a ring slot an EARLIER frame of the launch still has to read
				// (frame f reads slots written SIZE - t - 2 .. SIZE - t frames later once the launch is longer than SIZE - t)
				const float span = (float)std::min(n, 131072) + std::max(tl, tr) + 4.f;
				p.mode = (p.chunk >= cf + 2 && tl < 192000.f && tr < 192000.f && span < 192000.f) ? KB_PLAN_PARALLEL : KB_PLAN_SEQUENTIAL;
				p.gain = p.delay = p.dry = 0.f;
				all_parallel = all_parallel && p.mode == KB_PLAN_PARALLEL;
			}
			if (b->plan_cache.size() != plan.size() || memcmp(b->plan_cache.data(), plan.data(), plan.size() * sizeof(KbFxPlan)) != 0) {
				KB_CUDA(cudaMemcpyAsync(b->d_plan, plan.data(), plan.size() * sizeof(KbFxPlan), cudaMemcpyHostToDevice, b->stream));
				KB_CUDA(cudaStreamSynchronize(b->stream));
				b->plan_cache = plan;
			}
		}
		// streaming launches of at most 131072 frames (a longer launch could overwrite ring samples an earlier chunk still has to read)
		for (int o = 0; o < n; o += 131072) {
			const int len = std::min(131072, n - o);
			if (!seq_only) {
				// one launch: CTAs = chunks x instances, ordered by an in-kernel ticket, per-instance look-back on the delays
				KbDppSync* sync = (KbDppSync*)b->d_sync;
				const int chunks = (len + cf - 1) / cf;
				++b->epoch;
				if (cf == 512) kb_dpingpong_stream_kernel<512><<<chunks * b->instances, 256, 0, b->stream>>>(b->d_hdr, st, b->d_plan, b->d_rings, d + o, len, n, b->instances, b->fs, sync, b->epoch);
				else if (cf == 2048) kb_dpingpong_stream_kernel<2048><<<chunks * b->instances, 256, 0, b->stream>>>(b->d_hdr, st, b->d_plan, b->d_rings, d + o, len, n, b->instances, b->fs, sync, b->epoch);
				else kb_dpingpong_stream_kernel<1024><<<chunks * b->instances, 256, 0, b->stream>>>(b->d_hdr, st, b->d_plan, b->d_rings, d + o, len, n, b->instances, b->fs, sync, b->epoch);
				b->launches++;
			}
			if (!all_parallel) {
				kb_fx_seq_kernel<KB_FX_DELAY_PINGPONG, KbDPingPong><<<ib, 32, 0, b->stream>>>(b->d_hdr, st, b->d_rings, d + o, len, n, 2, b->instances, b->fs, seq_only ? nullptr : b->d_plan);
				if (o + 131072 < n || !seq_only) b->launches++;
			}
		}
		break; }
	case KB_FX_DELAY_REVERB: {
		KbDReverb* st = (KbDReverb*)b->d_state;
		if (!seq_only) {
			kb_dreverb_plan_kernel<<<ib, 32, 0, b->stream>>>(b->d_hdr, st, b->d_plan, b->instances, b->fs);
			kb_dreverb_par_kernel<<<b->instances, 512, 0, b->stream>>>(b->d_hdr, st, b->d_plan, b->d_rings, d, n, n, b->fs);
			b->launches += 2;
		}
		kb_fx_seq_kernel<KB_FX_DELAY_REVERB, KbDReverb><<<ib, 32, 0, b->stream>>>(b->d_hdr, st, b->d_rings, d, n, n, 1, b->instances, b->fs, seq_only ? nullptr : b->d_plan);
		break; }
	}
	b->prof_end();
	b->launches++;
	KB_CUDA(cudaGetLastError());
	if (b->graph != KB_FX_GAIN && b->graph != KB_FX_PAN && b->graph != KB_FX_CLIPPING && b->graph != KB_FX_FUNCTIONS && b->graph != KB_FX_MUTE) b->host_stale = true;
	if (!(flags & KB_DEVICE_PTR)) {
		KB_CUDA(cudaMemcpyAsync(io, d, floats * sizeof(float), cudaMemcpyDeviceToHost, b->stream));
		if (!(flags & KB_ASYNC_HOST)) KB_CUDA(cudaStreamSynchronize(b->stream));
	}
	return KB_OK;
}

// ================================================================================================ synths
struct kb_synth_bank : kb_bank_base {
	int graph = 0, instances = 0, voices = 0, channels = 1, ncontrols = 0;
	size_t voice_bytes = 0;
	std::vector<KbControl> controls;              // [instances][KB_MAX_CONTROLS], host-owned
	// host mirror in pinned memory (asynchronous, full-speed copies); dirty voices are uploaded packed
	KbVoiceHdr* hdr = nullptr; unsigned char* vstate = nullptr; size_t hdr_count = 0, vstate_bytes = 0;
	// pinned upload staging: a ring of KB_NSTAGE slots, each guarded by an event recorded after its H2D copies, so a caller
	// that runs ahead of the device (KB_DEVICE_PTR calls are asynchronous) never rewrites a slot whose copy is still queued
	static constexpr int KB_NSTAGE = 4;
	unsigned char* staging = nullptr;                                           // pinned, KB_NSTAGE slots
	cudaEvent_t stage_done[KB_NSTAGE] = {}; bool stage_used[KB_NSTAGE] = {}; int stage_next = 0;
	size_t stage_bytes = 0;
	unsigned char* d_staging = nullptr;
	// the staged records also travel to a device slot by an asynchronous copy on a stream of its own, issued when the host has packed them (the
	// GPU is still rendering the previous block), and the voice kernel gathers them from there: a CTA with a re-written voice no longer starts
	// its tile loop a PCIe round trip (~2.5 us) after the others.  KB_STAGE_COPY=0: the kernel reads the pinned slot itself (A/B, same results)
	unsigned char* d_stage_slots = nullptr; cudaStream_t stage_stream = nullptr; cudaEvent_t stage_copied[KB_NSTAGE] = {};
	std::vector<unsigned char> voice_dirty; std::vector<int> dirty_list; bool all_dirty = true;
	long long h2d_bytes = 0, d2h_bytes = 0;                                     // state traffic so far (for the e2e accounting)
	bool hdr_stale = false, vstate_stale = false;                               // finer than host_stale: what the device has changed
	// graphs whose Note::on() rewrites every field the device evolves: a start needs no state fetch (Filter.k:15-23, SuperSaw.k:12-19)
	bool on_overwrites() const { return graph == KB_SY_SUBTRACTIVE || graph == KB_SY_FILTER_K || graph == KB_SY_SUPERSAW; }
	std::vector<unsigned> noteOns, noteStart;     // Notes::noteOns / noteStart  klang.h:4333-4334
	void mark_dirty(int v) { dirty = true; if (!voice_dirty[v]) { voice_dirty[v] = 1; dirty_list.push_back(v); } }
	std::vector<KbSynthBlock> blk; bool blk_dirty = true;
	KbVoiceHdr* d_hdr = nullptr; unsigned char* d_vstate = nullptr; KbSynthBlock* d_blk = nullptr;
	float *d_scratch = nullptr, *d_out = nullptr, *d_adsr = nullptr, *d_mix = nullptr;
	// kb_synth_bank_process_mixdown: the exchange kernel of block k runs on a side stream beside the voice kernels of block k + 1
	cudaStream_t mix_stream = nullptr; cudaEvent_t ev_mix_in = nullptr; kb_mixdown* last_mixdown = nullptr;
	// fused mix-down: the per-instance sums are double-buffered (block k's exchange kernel, on the side stream, reads buffer k & 1 while block
	// k + 1's mix kernel writes the other one), so no event wait sits between a block's voice kernel and its mix kernel
	float* d_out_alt = nullptr; cudaEvent_t ev_exch[2] = { nullptr, nullptr }; unsigned exch_count = 0;
	KbStaged staged = {}; int staged_slot = -1;   // dirty-voice records the next voice kernel picks up itself (sy_upload with defer_scatter)
	int total() const { return instances * voices; }
	template <class T> T& vs(int v) { return *reinterpret_cast<T*>(vstate + (size_t)v * voice_bytes); }
	KbControl* ctl(int inst) { return controls.data() + (size_t)inst * KB_MAX_CONTROLS; }
};

static int sy_fetch(kb_synth_bank* b, bool need_hdr = true, bool need_vstate = true) {
	const bool get_hdr = need_hdr && b->hdr_stale, get_vs = need_vstate && b->vstate_stale;
	if (!get_hdr && !get_vs) return KB_OK;
	KB_CUDA(cudaSetDevice(b->device));
	if (get_vs && !b->dirty_list.empty()) {
		// voices re-written since the last block (their mirror copy is newer than the device's) must survive the fetch
		KB_CUDA(cudaStreamSynchronize(b->stream));
		std::vector<unsigned char> keep(b->dirty_list.size() * b->voice_bytes);
		for (size_t k = 0; k < b->dirty_list.size(); k++) memcpy(keep.data() + k * b->voice_bytes, b->vstate + (size_t)b->dirty_list[k] * b->voice_bytes, b->voice_bytes);
		KB_CUDA(cudaMemcpy(b->vstate, b->d_vstate, b->vstate_bytes, cudaMemcpyDeviceToHost));
		for (size_t k = 0; k < b->dirty_list.size(); k++) memcpy(b->vstate + (size_t)b->dirty_list[k] * b->voice_bytes, keep.data() + k * b->voice_bytes, b->voice_bytes);
		b->d2h_bytes += (long long)b->vstate_bytes;
		b->vstate_stale = false;
	} else if (get_vs) {
		KB_CUDA(cudaMemcpyAsync(b->vstate, b->d_vstate, b->vstate_bytes, cudaMemcpyDeviceToHost, b->stream));
		b->d2h_bytes += (long long)b->vstate_bytes;
	}
	if (get_hdr) {
		// (headers of re-started voices: stage/pitch/velocity were set by the host after the block, keep them)
		std::vector<KbVoiceHdr> keep;
		for (int v : b->dirty_list) keep.push_back(b->hdr[v]);
		KB_CUDA(cudaMemcpyAsync(b->hdr, b->d_hdr, b->hdr_count * sizeof(KbVoiceHdr), cudaMemcpyDeviceToHost, b->stream));
		KB_CUDA(cudaStreamSynchronize(b->stream));
		for (size_t k = 0; k < keep.size(); k++) b->hdr[b->dirty_list[k]] = keep[k];
		b->d2h_bytes += (long long)(b->hdr_count * sizeof(KbVoiceHdr));
		b->hdr_stale = false;
	}
	if (get_vs) { KB_CUDA(cudaStreamSynchronize(b->stream)); b->vstate_stale = false; }
	b->host_stale = b->hdr_stale || b->vstate_stale;
	return KB_OK;
}
static int sy_upload(kb_synth_bank* b, bool defer_scatter = false) {
	b->staged.records = nullptr; b->staged.count = 0; b->staged.voice_bytes = (int)b->voice_bytes;
	if (b->dirty) {
		const size_t rec = sizeof(KbVoiceHdr) + b->voice_bytes;
		// a full copy is only legal while the whole mirror is current; otherwise only the re-written voices may travel
		const bool mirror_current = !b->hdr_stale && !b->vstate_stale;
		if (!b->all_dirty && !(mirror_current && b->dirty_list.size() * 4 > b->hdr_count)) {
			const int count = (int)b->dirty_list.size();
			const int slot = b->stage_next; b->stage_next = (slot + 1) % kb_synth_bank::KB_NSTAGE;
			if (b->stage_used[slot]) KB_CUDA(cudaEventSynchronize(b->stage_done[slot]));    // its previous copies have left the host buffer
			// one slot = [voice index list, padded to 16 bytes][packed records]: ONE host-to-device copy per block
			unsigned char* stage = b->staging + (size_t)slot * b->stage_bytes;
			const size_t rec_off = ((size_t)count * sizeof(int) + 15) & ~(size_t)15;
			int* stage_index = (int*)stage;
			for (int k = 0; k < count; k++) {
				const int v = b->dirty_list[k];
				stage_index[k] = v;
				memcpy(stage + rec_off + (size_t)k * rec, b->hdr + v, sizeof(KbVoiceHdr));
				memcpy(stage + rec_off + (size_t)k * rec + sizeof(KbVoiceHdr), b->vstate + (size_t)v * b->voice_bytes, b->voice_bytes);
			}
			if (defer_scatter && count <= KB_STAGED_MAX) {
				// the voice kernel reads the records from this pinned slot itself (kb_tile_scatter); the slot is free again when that kernel has
				// run: sy_process records stage_done[slot] behind it
				b->staged.records = stage + rec_off; b->staged.count = count;
				static const bool stage_copy = !getenv("KB_STAGE_COPY") || atoi(getenv("KB_STAGE_COPY")) != 0;
				if (stage_copy) {
					if (!b->d_stage_slots) {
						KB_CUDA(cudaMalloc(&b->d_stage_slots, kb_synth_bank::KB_NSTAGE * b->stage_bytes));
						KB_CUDA(cudaStreamCreateWithFlags(&b->stage_stream, cudaStreamNonBlocking));
						for (int k = 0; k < kb_synth_bank::KB_NSTAGE; k++) KB_CUDA(cudaEventCreateWithFlags(&b->stage_copied[k], cudaEventDisableTiming));
					}
					unsigned char* dslot = b->d_stage_slots + (size_t)slot * b->stage_bytes;
					KB_CUDA(cudaMemcpyAsync(dslot + rec_off, stage + rec_off, (size_t)count * rec, cudaMemcpyHostToDevice, b->stage_stream));
					KB_CUDA(cudaEventRecord(b->stage_copied[slot], b->stage_stream));
					KB_CUDA(cudaStreamWaitEvent(b->stream, b->stage_copied[slot], 0));
					b->staged.records = dslot + rec_off;
				}
				for (int k = 0; k < count; k++) b->staged.index[k] = stage_index[k];
				b->staged_slot = slot;
			} else {
				KB_CUDA(cudaMemcpyAsync(b->d_staging, stage, rec_off + (size_t)count * rec, cudaMemcpyHostToDevice, b->stream));
				KB_CUDA(cudaEventRecord(b->stage_done[slot], b->stream));
				const int words = count * (int)(rec / 4);
				kb_scatter_voices_kernel<<<std::min(148, (words + 255) / 256), 256, 0, b->stream>>>(b->d_staging + rec_off, (const int*)b->d_staging, count, (int)b->voice_bytes, b->d_hdr, b->d_vstate);
				b->launches++;
			}
			b->stage_used[slot] = true;
			b->h2d_bytes += (long long)count * (long long)(rec + sizeof(int));
		} else {
			KB_CUDA(cudaMemcpyAsync(b->d_hdr, b->hdr, b->hdr_count * sizeof(KbVoiceHdr), cudaMemcpyHostToDevice, b->stream));
			KB_CUDA(cudaMemcpyAsync(b->d_vstate, b->vstate, b->vstate_bytes, cudaMemcpyHostToDevice, b->stream));
			KB_CUDA(cudaStreamSynchronize(b->stream));      // the source is the live mirror, which the next event rewrites (rare path: first upload / most voices dirty)
			b->h2d_bytes += (long long)(b->hdr_count * sizeof(KbVoiceHdr) + b->vstate_bytes);
		}
		for (int v : b->dirty_list) b->voice_dirty[v] = 0;
		b->dirty_list.clear(); b->all_dirty = false;
	}
	if (b->blk_dirty) {
		for (int i = 0; i < b->instances; i++) {
			KbSynthBlock& k = b->blk[i];
			memset(&k, 0, sizeof(k));
			if (b->graph == KB_SY_TB303) k.tb = kb_tb_block(b->fs, b->ctl(i));
			if (b->graph == KB_SY_SYNTHX) { k.sx_tr_at = kb_sx_tr_at(b->ctl(i)[2].value); k.sx_dt_at = kb_sx_dt_at(b->ctl(i)[1].value); }
			if (b->graph == KB_SY_FM) { k.fm_i1 = b->ctl(i)[1].value; k.fm_i2 = b->ctl(i)[2].value; }
			if (b->graph >= KB_SY_AM && b->graph <= KB_SY_MOD_FM2) for (int c = 0; c < 3; c++) k.c[c] = b->ctl(i)[c].value;   // (graph 14 has no controls)
			else k.c[0] = k.c[1] = k.c[2] = 0.f;
		}
		KB_CUDA(cudaMemcpyAsync(b->d_blk, b->blk.data(), b->blk.size() * sizeof(KbSynthBlock), cudaMemcpyHostToDevice, b->stream));
		KB_CUDA(cudaStreamSynchronize(b->stream));      // blk is pageable and may be rewritten by the next control change
	}
	b->dirty = false; b->blk_dirty = false;
	return KB_OK;
}

extern "C" kb_synth_bank* kb_synth_bank_create(int graph, int instances, int voices, float fs, int max_block, int device) {
	if (graph < 0 || graph >= KB_SY_COUNT || instances < 1 || voices < 1 || voices > KB_MAX_VOICES || max_block < 1 || !(fs > 0)) {
		kb_fail(KB_EINVAL, "kb_synth_bank_create: bad argument"); return nullptr;
	}
	// (kernels carry the instance index — the additive graphs the voice index — in gridDim.y, which is limited to 65535)
	if (instances > 65535 || ((graph == KB_SY_ADDITIVE_SAW || graph == KB_SY_ADDITIVE_SQUARE || graph == KB_SY_ADDITIVE_NYQUIST) && (long long)instances * std::max(voices, 32) > 65535)) {
		kb_fail(KB_EINVAL, "kb_synth_bank_create: at most 65535 instances (additive graphs: 65535 voices) per bank (create several banks)"); return nullptr;
	}
	if (device < 0 || device >= kb_device_count()) { kb_fail(KB_ENODEV, "kb_synth_bank_create: no such CUDA device (klang-b200 has no CPU path)"); return nullptr; }
	kb_synth_bank* b = new kb_synth_bank();
	b->graph = graph; b->instances = instances; b->device = device; b->max_block = max_block; b->fs = kb_make_fs(fs);
	// the example synths allocate at least 32 notes (Filter.k:39, SuperSaw.k:52, TB303.k:128, SynTHX.k:198)
	if (graph != KB_SY_SUBTRACTIVE && voices < 32) voices = 32;
	b->voices = voices;
	b->controls.assign((size_t)instances * KB_MAX_CONTROLS, KbControl{ 0, 0, 0, 0 });
	switch (graph) {
	case KB_SY_SUBTRACTIVE: b->ncontrols = 4; b->voice_bytes = sizeof(KbSubVoice); break;
	case KB_SY_FILTER_K: b->ncontrols = 0; b->voice_bytes = sizeof(KbSubVoice); break;
	case KB_SY_SUPERSAW: b->ncontrols = 3; b->voice_bytes = sizeof(KbSsawVoice); break;
	case KB_SY_TB303: b->ncontrols = 5; b->voice_bytes = sizeof(KbTbVoice); break;
	case KB_SY_SYNTHX: b->ncontrols = 5; b->voice_bytes = sizeof(KbSxVoice); b->channels = 2; break;
	case KB_SY_FM: b->ncontrols = 4; b->voice_bytes = sizeof(KbFmVoice); break;
	case KB_SY_BREAKPOINT: b->ncontrols = 2; b->voice_bytes = sizeof(KbSenvVoice); break;
	case KB_SY_RAMP: b->ncontrols = 1; b->voice_bytes = sizeof(KbSenvVoice); break;
	case KB_SY_RELEASE: b->ncontrols = 4; b->voice_bytes = sizeof(KbSenvVoice); break;
	case KB_SY_ADDITIVE_SAW: case KB_SY_ADDITIVE_SQUARE: case KB_SY_ADDITIVE_NYQUIST: b->ncontrols = 0; b->voice_bytes = sizeof(KbAddVoice); break;
	case KB_SY_AM: case KB_SY_MOD_FM: b->ncontrols = 2; b->voice_bytes = sizeof(KbSmodVoice); break;
	case KB_SY_MOD_FM2: b->ncontrols = 3; b->voice_bytes = sizeof(KbSmodVoice); break;
	}
	for (int i = 0; i < instances; i++) {
		KbControl* c = b->ctl(i);
		switch (graph) {
		case KB_SY_SUBTRACTIVE: c[0] = kb_dial(0.f, 1.f, 0.01f); c[1] = kb_dial(0.f, 1.f, 0.1f); c[2] = kb_dial(0.f, 1.f, 0.7f); c[3] = kb_dial(0.f, 1.f, 0.25f); break;
		case KB_SY_SUPERSAW: c[0] = kb_dial(0.001f, 1.f, 0.001f); c[1] = kb_dial(0.f, 1.f, 0.05f); c[2] = kb_dial(0.f, 1.f, 0.6f); break;   // SuperSaw.k:38-43
		case KB_SY_TB303: c[0] = kb_dial(0.f, 1.f, 1.f); c[1] = kb_dial(0.f, 1.f, 0.5f); c[2] = kb_dial(0.1f, 1.f, 0.5f);                     // TB303.k:118-126
		                  c[3] = kb_dial(0.f, 1.f, 0.f); c[4] = kb_dial(0.01f, 10.f, 1.f); break;
		case KB_SY_FM: c[0] = kb_dial(0.001f, 10.f, 1.0f); c[1] = kb_dial(0.f, 10.f, 0.37f); c[2] = kb_dial(0.f, 10.f, 0.37f); c[3] = kb_dial(0.f, 1.f, 0.5f); break;   // FM.k:80-86
		case KB_SY_BREAKPOINT: c[0] = kb_dial(0.05f, 1.0f, 0.05f); c[1] = kb_dial(0.1f, 1.0f, 0.1f); break;                                  // Breakpoint.k:25-28
		case KB_SY_RAMP: c[0] = kb_dial(0.1f, 1.0f, 0.1f); break;                                                                            // Ramp.k:23-25
		case KB_SY_RELEASE: c[0] = kb_dial(0.f, 1.f, 0.002f); c[1] = kb_dial(0.f, 1.f, 0.1f); c[2] = kb_dial(0.f, 1.f, 0.05f); c[3] = kb_dial(0.f, 1.f, 1.0f); break;   // Release.k:33-38
		case KB_SY_AM: c[0] = kb_dial(0.01f, 3.f, 0.5f); c[1] = kb_dial(0.f, 1.f, 0.5f); break;                                              // AM.k:33-36
		case KB_SY_MOD_FM: c[0] = kb_dial(0.01f, 3.f, 0.5f); c[1] = kb_dial(0.f, 10.f, 0.5f); break;                                         // Modulation/FM.k:37-40
		case KB_SY_MOD_FM2: c[0] = kb_dial(0.01f, 3.f, 0.5f); c[1] = kb_dial(0.f, 10.f, 0.5f); c[2] = kb_dial(0.f, 10.f, 0.5f); break;       // FM2.k:38-42
		case KB_SY_SYNTHX: c[0] = kb_dial(0.f, 5.f, 0.5f); c[1] = kb_dial(0.f, 1.f, 0.5f); c[2] = kb_dial(0.f, 1.f, 0.6f);                   // SynTHX.k:186-194
		                   c[3] = kb_dial(0.f, 1.f, 1.f); c[4] = kb_dial(0.f, 1.f, 0.f); break;
		}
	}
	const int total = b->total();
	b->hdr_count = total; b->vstate_bytes = (size_t)total * b->voice_bytes;
	if (cudaSetDevice(device) != cudaSuccess || cudaHostAlloc((void**)&b->hdr, total * sizeof(KbVoiceHdr), cudaHostAllocDefault) != cudaSuccess ||
	    cudaHostAlloc((void**)&b->vstate, b->vstate_bytes, cudaHostAllocDefault) != cudaSuccess ||
	    cudaHostAlloc((void**)&b->staging, kb_synth_bank::KB_NSTAGE * ((size_t)(total + 1) * (sizeof(KbVoiceHdr) + b->voice_bytes + sizeof(int)) + 16), cudaHostAllocMapped) != cudaSuccess) {
		kb_fail(KB_ECUDA, std::string("kb_synth_bank_create: pinned host allocation: ") + cudaGetErrorString(cudaGetLastError()));
		kb_synth_bank_destroy(b); return nullptr;
	}
	b->stage_bytes = (size_t)(total + 1) * (sizeof(KbVoiceHdr) + b->voice_bytes + sizeof(int)) + 16;
	for (int k = 0; k < kb_synth_bank::KB_NSTAGE; k++)
		if (cudaEventCreateWithFlags(&b->stage_done[k], cudaEventDisableTiming) != cudaSuccess) { kb_fail(KB_ECUDA, "kb_synth_bank_create: event"); kb_synth_bank_destroy(b); return nullptr; }
	for (int v = 0; v < total; v++) b->hdr[v] = KbVoiceHdr{ KB_NOTE_OFF, 0.f, 0.f, 0 };
	memset(b->vstate, 0, b->vstate_bytes);
	b->voice_dirty.assign(total, 0);
	b->noteOns.assign(instances, 0u); b->noteStart.assign(total, 0u);
	b->blk.assign(instances, KbSynthBlock());
	for (int v = 0; v < total; v++) {
		switch (graph) {
		case KB_SY_SUBTRACTIVE: case KB_SY_FILTER_K: kb_sub_construct(b->fs, graph, b->vs<KbSubVoice>(v)); break;
		case KB_SY_SUPERSAW: kb_ssaw_construct(b->fs, b->vs<KbSsawVoice>(v)); break;
		case KB_SY_FM: kb_fm_construct(b->fs, b->vs<KbFmVoice>(v)); break;
		case KB_SY_BREAKPOINT: case KB_SY_RAMP: case KB_SY_RELEASE: kb_senv_construct(b->fs, graph, b->vs<KbSenvVoice>(v)); break;
		case KB_SY_ADDITIVE_SAW: case KB_SY_ADDITIVE_SQUARE: case KB_SY_ADDITIVE_NYQUIST: kb_add_construct(graph, b->vs<KbAddVoice>(v)); break;
		case KB_SY_AM: case KB_SY_MOD_FM: case KB_SY_MOD_FM2: kb_smod_construct(b->fs, graph, b->vs<KbSmodVoice>(v)); break;
		case KB_SY_TB303: kb_tb_construct(b->fs, b->vs<KbTbVoice>(v)); break;
		case KB_SY_SYNTHX: kb_sx_construct(b->fs, b->vs<KbSxVoice>(v)); break;
		}
	}
	bool ok = cudaSetDevice(device) == cudaSuccess;
	ok = ok && cudaStreamCreateWithFlags(&b->own_stream, cudaStreamNonBlocking) == cudaSuccess;
	b->stream = b->own_stream;
	ok = ok && dev_alloc(&b->d_hdr, total) == cudaSuccess;
	ok = ok && dev_alloc(&b->d_vstate, b->vstate_bytes) == cudaSuccess;
	ok = ok && dev_alloc(&b->d_staging, (size_t)(total + 1) * (sizeof(KbVoiceHdr) + b->voice_bytes + sizeof(int)) + 16) == cudaSuccess;
	ok = ok && dev_alloc(&b->d_blk, instances) == cudaSuccess;
	ok = ok && dev_alloc(&b->d_scratch, (size_t)total * b->channels * max_block) == cudaSuccess;
	ok = ok && dev_alloc(&b->d_out, (size_t)instances * b->channels * max_block) == cudaSuccess;
	ok = ok && dev_alloc(&b->d_mix, (size_t)b->channels * max_block) == cudaSuccess;
	if (graph == KB_SY_SYNTHX) ok = ok && dev_alloc(&b->d_adsr, (size_t)total * max_block) == cudaSuccess;
	if (graph == KB_SY_SYNTHX) ok = ok && cudaFuncSetAttribute(kb_sx_render_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(KbSxSmem)) == cudaSuccess;
	if (!ok) { kb_fail(KB_ECUDA, std::string("kb_synth_bank_create: ") + cudaGetErrorString(cudaGetLastError())); kb_synth_bank_destroy(b); return nullptr; }
	return b;
}
extern "C" void kb_synth_bank_destroy(kb_synth_bank* b) {
	if (!b) return;
	cudaSetDevice(b->device);
	if (b->stream) cudaStreamSynchronize(b->stream);
	b->prof_free();
	for (int k = 0; k < kb_synth_bank::KB_NSTAGE; k++) if (b->stage_done[k]) cudaEventDestroy(b->stage_done[k]);
	cudaFreeHost(b->hdr); cudaFreeHost(b->vstate); cudaFreeHost(b->staging);
	cudaFree(b->d_staging);
	if (b->stage_stream) { cudaStreamSynchronize(b->stage_stream); cudaStreamDestroy(b->stage_stream); }
	for (int k = 0; k < kb_synth_bank::KB_NSTAGE; k++) if (b->stage_copied[k]) cudaEventDestroy(b->stage_copied[k]);
	cudaFree(b->d_stage_slots);
	cudaFree(b->d_hdr); cudaFree(b->d_vstate); cudaFree(b->d_blk); cudaFree(b->d_scratch); cudaFree(b->d_out); cudaFree(b->d_adsr); cudaFree(b->d_mix);
	if (b->mix_stream) { cudaStreamSynchronize(b->mix_stream); cudaStreamDestroy(b->mix_stream); }
	if (b->ev_mix_in) cudaEventDestroy(b->ev_mix_in);
	for (int k = 0; k < 2; k++) if (b->ev_exch[k]) cudaEventDestroy(b->ev_exch[k]);
	cudaFree(b->d_out_alt);
	if (b->own_stream) cudaStreamDestroy(b->own_stream);
	delete b;
}
extern "C" int kb_synth_bank_channels(const kb_synth_bank* b) { return b ? b->channels : KB_EINVAL; }
extern "C" int kb_synth_bank_instances(const kb_synth_bank* b) { return b ? b->instances : KB_EINVAL; }
extern "C" int kb_synth_bank_voices(const kb_synth_bank* b) { return b ? b->voices : KB_EINVAL; }
extern "C" int kb_synth_bank_num_controls(const kb_synth_bank* b) { return b ? b->ncontrols : KB_EINVAL; }
extern "C" long long kb_synth_bank_launches(const kb_synth_bank* b) { return b ? b->launches : 0; }
extern "C" long long kb_synth_bank_state_bytes(const kb_synth_bank* b) { return b ? (long long)(b->hdr_count * sizeof(KbVoiceHdr) + b->vstate_bytes) : 0; }
extern "C" int kb_synth_bank_transfer_bytes(const kb_synth_bank* b, long long* h2d, long long* d2h) { if (!b) return kb_fail(KB_EINVAL, "null bank"); if (h2d) *h2d = b->h2d_bytes; if (d2h) *d2h = b->d2h_bytes; return KB_OK; }
extern "C" int kb_synth_bank_profile(kb_synth_bank* b, int enable) { if (!b) return kb_fail(KB_EINVAL, "null bank"); cudaSetDevice(b->device); cudaStreamSynchronize(b->stream); b->profiling = enable != 0; b->prof_used = 0; return KB_OK; }
extern "C" int kb_synth_bank_profile_read(kb_synth_bank* b, double* ms, long long* count) { if (!b) return kb_fail(KB_EINVAL, "null bank"); cudaSetDevice(b->device); return b->prof_read(ms, count); }
extern "C" int kb_synth_bank_sync(kb_synth_bank* b) {
	if (!b) return kb_fail(KB_EINVAL, "null bank");
	KB_CUDA(cudaSetDevice(b->device)); KB_CUDA(cudaStreamSynchronize(b->stream));
	if (b->mix_stream) KB_CUDA(cudaStreamSynchronize(b->mix_stream));
	return KB_OK;
}
extern "C" int kb_synth_bank_set_stream(kb_synth_bank* b, void* s) {
	if (!b) return kb_fail(KB_EINVAL, "null bank");
	KB_CUDA(cudaStreamSynchronize(b->stream));
	b->stream = s ? (cudaStream_t)s : b->own_stream;
	return KB_OK;
}
extern "C" int kb_synth_bank_set_control(kb_synth_bank* b, int inst, int idx, float v) {
	if (!b || inst < 0 || inst >= b->instances || idx < 0 || idx >= b->ncontrols) return kb_fail(KB_EINVAL, "kb_synth_bank_set_control: bad argument");
	kb_control_set(b->ctl(inst)[idx], v);
	b->blk_dirty = true;
	return KB_OK;
}
extern "C" int kb_synth_bank_get_control(kb_synth_bank* b, int inst, int idx, float* v) {
	if (!b || !v || inst < 0 || inst >= b->instances || idx < 0 || idx >= b->ncontrols) return kb_fail(KB_EINVAL, "kb_synth_bank_get_control: bad argument");
	*v = b->ctl(inst)[idx].value;
	return KB_OK;
}

// Synth::onControl / onPreset (klang.h:4399-4404, 4415-4420; stereo 4789-4794, 4805-4810): the synth's own control() / preset() hook, then the
// hook of every note whose stage is not Off.  No bound program overrides either hook (their process() reads controls[] directly), so the call
// has no effect on the audio; what is observable — and returned — is how many notes the reference would notify, which needs the stages the
// DEVICE has evolved (a note that finished inside a block is Off), fetched here when stale.
static int sy_notified(kb_synth_bank* b, int inst) {
	int rc = sy_fetch(b, true, false); if (rc) return rc;
	int n = 0;
	for (int v = 0; v < b->voices; v++) n += b->hdr[(size_t)inst * b->voices + v].stage != KB_NOTE_OFF;
	return n;
}
extern "C" int kb_synth_bank_on_control(kb_synth_bank* b, int inst, int idx, float value) {
	(void)value;
	if (!b || inst < 0 || inst >= b->instances || idx < 0) return kb_fail(KB_EINVAL, "kb_synth_bank_on_control: bad argument");
	return sy_notified(b, inst);
}
extern "C" int kb_synth_bank_on_preset(kb_synth_bank* b, int inst, int index) {
	if (!b || inst < 0 || inst >= b->instances || index < 0) return kb_fail(KB_EINVAL, "kb_synth_bank_on_preset: bad argument");
	return sy_notified(b, inst);
}
extern "C" int kb_synth_bank_load_preset(kb_synth_bank* b, int inst, int index) {
	if (!b || inst < 0 || inst >= b->instances) return kb_fail(KB_EINVAL, "kb_synth_bank_load_preset: bad argument");
	const KbPreset* p = index >= 0 ? kb_preset_find(1, b->graph, index) : nullptr;
	if (!p) return kb_fail(KB_EINVAL, "kb_synth_bank_load_preset: the bank's program has no such preset");
	for (int c = 0; c < p->count && c < b->ncontrols; c++) { int rc = kb_synth_bank_set_control(b, inst, c, p->values[c]); if (rc) return rc; }
	return kb_synth_bank_on_preset(b, inst, index);
}

// NoteBase::start: stage = Onset; on(pitch, velocity); stage = Sustain           klang.h:4257-4263
static void sy_start(kb_synth_bank* b, int inst, int voice, float pitch, float velocity) {
	const int v = inst * b->voices + voice;
	KbVoiceHdr& h = b->hdr[v];
	h.pitch = pitch; h.velocity = velocity;
	const KbControl* c = b->ctl(inst);
	switch (b->graph) {
	case KB_SY_SUBTRACTIVE: case KB_SY_FILTER_K: kb_sub_on(b->fs, b->graph, c, b->vs<KbSubVoice>(v), pitch); break;
	case KB_SY_SUPERSAW: kb_ssaw_on(b->fs, c, b->vs<KbSsawVoice>(v), pitch); break;
	case KB_SY_FM: kb_fm_on(b->fs, c, b->vs<KbFmVoice>(v), pitch); break;
	case KB_SY_BREAKPOINT: case KB_SY_RAMP: case KB_SY_RELEASE: kb_senv_on(b->fs, b->graph, c, b->vs<KbSenvVoice>(v), pitch); break;
	case KB_SY_ADDITIVE_SAW: case KB_SY_ADDITIVE_SQUARE: case KB_SY_ADDITIVE_NYQUIST: kb_add_on(b->fs, b->vs<KbAddVoice>(v), pitch); break;
	case KB_SY_AM: case KB_SY_MOD_FM: case KB_SY_MOD_FM2: kb_smod_on(b->fs, b->vs<KbSmodVoice>(v), pitch); break;
	case KB_SY_TB303: kb_tb_on(b->fs, c, b->vs<KbTbVoice>(v), pitch); break;
	case KB_SY_SYNTHX: kb_sx_on(b->fs, c, b->vs<KbSxVoice>(v), pitch); break;
	}
	h.stage = KB_NOTE_SUSTAIN;
	b->mark_dirty(v);
}
// NoteBase::release: Off stays Off; otherwise stage = Release; off(velocity)     klang.h:4265-4275
static void sy_release(kb_synth_bank* b, int inst, int voice) {
	const int v = inst * b->voices + voice;
	KbVoiceHdr& h = b->hdr[v];
	if (h.stage == KB_NOTE_OFF || h.stage == KB_NOTE_RELEASE) return;
	h.stage = KB_NOTE_RELEASE;
	switch (b->graph) {
	case KB_SY_SUBTRACTIVE: case KB_SY_FILTER_K: kb_adsr_release(b->fs, b->vs<KbSubVoice>(v).adsr); break;   // Filter.k:25-27
	case KB_SY_SUPERSAW: kb_adsr_release(b->fs, b->vs<KbSsawVoice>(v).adsr); break;                          // SuperSaw.k:21-23
	case KB_SY_FM: kb_adsr_release(b->fs, b->vs<KbFmVoice>(v).adsr); break;                                  // FM.k:56-58
	case KB_SY_RELEASE: kb_env_release(b->fs, b->vs<KbSenvVoice>(v).env, b->ctl(inst)[3].value, 0.f); break;  // Release.k:21-24
	case KB_SY_AM: case KB_SY_MOD_FM: case KB_SY_MOD_FM2: kb_adsr_release(b->fs, b->vs<KbSmodVoice>(v).adsr); break;   // AM.k:17-19
	case KB_SY_BREAKPOINT: case KB_SY_RAMP: case KB_SY_ADDITIVE_SAW: case KB_SY_ADDITIVE_SQUARE: case KB_SY_ADDITIVE_NYQUIST:
		h.stage = KB_NOTE_OFF; break;                                                                        // NoteBase::off default: stage = Off  klang.h:4237
	case KB_SY_TB303: kb_adsr_release(b->fs, b->vs<KbTbVoice>(v).adsr); break;                               // TB303.k:99-101
	case KB_SY_SYNTHX: kb_adsr_release(b->fs, b->vs<KbSxVoice>(v).adsr); break;                              // SynTHX.k:163-165
	}
	b->mark_dirty(v);
}
// Notes::assign: first Off voice, else the oldest Released, else the oldest      klang.h:4336-4372
static int sy_assign(kb_synth_bank* b, int inst) {
	const KbVoiceHdr* h = b->hdr + (size_t)inst * b->voices;
	unsigned* start = b->noteStart.data() + (size_t)inst * b->voices;
	unsigned& ons = b->noteOns[inst];
	for (int i = 0; i < b->voices; i++) if (h[i].stage == KB_NOTE_OFF) { start[i] = ons++; return i; }
	int oldest = -1; unsigned oldest_start = 0;
	for (int i = 0; i < b->voices; i++)
		if (h[i].stage == KB_NOTE_RELEASE && (oldest == -1 || start[i] < oldest_start)) { oldest = i; oldest_start = start[i]; }
	if (oldest != -1) { start[oldest] = ons++; return oldest; }
	for (int i = 0; i < b->voices; i++) if (oldest == -1 || start[i] < oldest_start) { oldest = i; oldest_start = start[i]; }
	start[oldest] = ons++;
	return oldest;
}
extern "C" int kb_synth_bank_note_on(kb_synth_bank* b, int inst, int pitch, float velocity) {
	if (!b || inst < 0 || inst >= b->instances) return kb_fail(KB_EINVAL, "kb_synth_bank_note_on: bad argument");
	int rc = sy_fetch(b, true, !b->on_overwrites()); if (rc) return rc;
	const int n = sy_assign(b, inst);
	sy_start(b, inst, n, (float)pitch, velocity);
	return n;
}
extern "C" int kb_synth_bank_note_off(kb_synth_bank* b, int inst, int pitch, float velocity) {
	(void)velocity;
	if (!b || inst < 0 || inst >= b->instances) return kb_fail(KB_EINVAL, "kb_synth_bank_note_off: bad argument");
	int rc = sy_fetch(b); if (rc) return rc;
	for (int n = 0; n < b->voices; n++) {
		const KbVoiceHdr& h = b->hdr[(size_t)inst * b->voices + n];
		if (h.pitch == pitch && h.stage == KB_NOTE_SUSTAIN) sy_release(b, inst, n);
	}
	return KB_OK;
}
extern "C" int kb_synth_bank_voice_start(kb_synth_bank* b, int inst, int voice, float pitch, float velocity) {
	if (!b || inst < 0 || inst >= b->instances || voice < 0 || voice >= b->voices) return kb_fail(KB_EINVAL, "kb_synth_bank_voice_start: bad argument");
	int rc = sy_fetch(b, false, !b->on_overwrites()); if (rc) return rc;
	sy_start(b, inst, voice, pitch, velocity);
	return KB_OK;
}
extern "C" int kb_synth_bank_voice_release(kb_synth_bank* b, int inst, int voice, float velocity) {
	(void)velocity;
	if (!b || inst < 0 || inst >= b->instances || voice < 0 || voice >= b->voices) return kb_fail(KB_EINVAL, "kb_synth_bank_voice_release: bad argument");
	int rc = sy_fetch(b); if (rc) return rc;
	sy_release(b, inst, voice);
	return KB_OK;
}
extern "C" int kb_synth_bank_voice_stage(kb_synth_bank* b, int inst, int voice) {
	if (!b || inst < 0 || inst >= b->instances || voice < 0 || voice >= b->voices) return kb_fail(KB_EINVAL, "kb_synth_bank_voice_stage: bad argument");
	int rc = sy_fetch(b, true, false); if (rc) return rc;
	return b->hdr[(size_t)inst * b->voices + voice].stage;
}

// Synth::input(status, byte1, byte2) of the v0.7.2 template (templates/juce/synth/Source/klang.h:3921-3929): the status byte is compared
// whole (channel 1 only, as there); every other message goes to onMIDI(), which none of the bound graphs overrides.
extern "C" int kb_synth_bank_midi(kb_synth_bank* b, int inst, int status, int byte1, int byte2) {
	if (!b || inst < 0 || inst >= b->instances) return kb_fail(KB_EINVAL, "kb_synth_bank_midi: bad argument");
	if (status == 0x90 && byte2 > 0) { const int rc = kb_synth_bank_note_on(b, inst, byte1, byte2 / 127.f); return rc > 0 ? KB_OK : rc; }
	if (status == 0x80 || (status == 0x90 && byte2 == 0)) return kb_synth_bank_note_off(b, inst, byte1, byte2 / 127.f);
	return KB_OK;
}

extern "C" int kb_synth_bank_events(kb_synth_bank* b, int count, const kb_note_event* ev) {
	if (!b || count < 0 || (count && !ev)) return kb_fail(KB_EINVAL, "kb_synth_bank_events: bad argument");
	for (int i = 0; i < count; i++) {
		int rc = KB_EINVAL;
		switch (ev[i].type) {
		case KB_EV_NOTE_ON: rc = kb_synth_bank_note_on(b, ev[i].instance, ev[i].key, ev[i].velocity); if (rc > 0) rc = 0; break;
		case KB_EV_NOTE_OFF: rc = kb_synth_bank_note_off(b, ev[i].instance, ev[i].key, ev[i].velocity); break;
		case KB_EV_VOICE_START: rc = kb_synth_bank_voice_start(b, ev[i].instance, ev[i].key, ev[i].pitch, ev[i].velocity); break;
		case KB_EV_VOICE_RELEASE: rc = kb_synth_bank_voice_release(b, ev[i].instance, ev[i].key, ev[i].velocity); break;
		case KB_EV_CONTROL: rc = kb_synth_bank_set_control(b, ev[i].instance, ev[i].key, ev[i].velocity); break;
		default: return kb_fail(KB_EINVAL, "kb_synth_bank_events: unknown event type");
		}
		if (rc < 0) return rc;
	}
	return KB_OK;
}

// is `p` page-locked host memory the device can address through the same pointer (cudaHostAlloc / cudaHostRegister under unified addressing)?
// Asked on every call: an address can be freed and come back as pageable memory.
static bool kb_host_ptr_mapped(const void* p) {
	cudaPointerAttributes a;
	if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
	return a.type == cudaMemoryTypeHost && a.devicePointer == p;
}
static int mixdown_step(kb_mixdown* m, const float* src, int rows, size_t row_stride, int count, float* out_prev, cudaStream_t stream);
static int sy_process(kb_synth_bank* b, float* out, int n, unsigned flags, kb_mixdown* mixdown, float* out_prev);
extern "C" int kb_synth_bank_process(kb_synth_bank* b, float* out, int n, unsigned flags) {
	if (!b || !out || n < 0 || n > b->max_block) return kb_fail(KB_EINVAL, "kb_synth_bank_process: bad argument (n > max_block?)");
	return sy_process(b, out, n, flags, nullptr, nullptr);
}
// events of the block, then the block: what a host's audio callback does (templates/juce/synth/Source/PluginProcessor.cpp:169-177), one call
extern "C" int kb_synth_bank_step(kb_synth_bank* b, int count, const kb_note_event* events, float* out, int n, unsigned flags) {
	int rc = kb_synth_bank_events(b, count, events); if (rc) return rc;
	return kb_synth_bank_process(b, out, n, flags);
}
static int sy_process(kb_synth_bank* b, float* out, int n, unsigned flags, kb_mixdown* mixdown, float* out_prev) {
	if (n == 0) return KB_OK;
	KB_CUDA(cudaSetDevice(b->device));
	const int total = b->total(), C = b->channels;
	// Subtractive / Filter.k, 800..1184 voices: the decoupled kernel with mbarrier hand-over (kb_sub_mbar_kernel, layout 4), which also scatters
	// the block's re-written voices itself.  KB_TILE_LAYOUT = 3 selects the polled-counter form (kb_sub_flow_kernel), 2 / 0 the lock-step
	// kernels, KB_TILE_G the voices per CTA (A/B measurement, same results)
	static const int force_g = getenv("KB_TILE_G") ? atoi(getenv("KB_TILE_G")) : 0;
	static const int layout = getenv("KB_TILE_LAYOUT") ? atoi(getenv("KB_TILE_LAYOUT")) : 4;
	const bool sub = b->graph == KB_SY_SUBTRACTIVE || b->graph == KB_SY_FILTER_K;
	int sub_g = total >= 1600 ? 16 : total > 7 * 148 ? 8 : total >= 800 ? 7 : 4;
	if (force_g) sub_g = force_g;
	const bool sub_flow = sub && !(flags & KB_LANE_PER_VOICE) && layout >= 3 && (sub_g == 7 || sub_g == 8);
	static const bool scatter_fused = !getenv("KB_SCATTER_FUSED") || atoi(getenv("KB_SCATTER_FUSED")) != 0;     // (A/B measurement)
	int rc = sy_upload(b, sub_flow && scatter_fused); if (rc) return rc;
	const bool per_voice = flags & KB_PER_VOICE, dev = flags & KB_DEVICE_PTR, bank_mix = (flags & KB_BANK_MIX) && !per_voice;
	const size_t out_floats = per_voice ? (size_t)total * C * n : bank_mix ? (size_t)C * n : (size_t)b->instances * C * n;
	cudaStream_t st = b->stream;
	// the exchange kernel of the previous block (side stream) reads d_out: the kernels that rewrite d_out wait for it (kb_mix_kernel, which is
	// queued behind this block's voice kernel; SynTHX's render kernel writes d_out itself, so it waits at once)
	auto wait_prev_exchange = [&]() -> int {
		if (b->last_mixdown && b->last_mixdown->step_pending) KB_CUDA(cudaStreamWaitEvent(st, b->last_mixdown->ev_done, 0));
		return KB_OK;
	};
	if (b->graph == KB_SY_SYNTHX || per_voice) { rc = wait_prev_exchange(); if (rc) return rc; }
	float* d_voice_dst = (per_voice && dev) ? out : b->d_scratch;
	float* d_result_fused = nullptr;                 // set when kb_mix_fused_kernel has also written the bank mix
	float* d_inst_dst = (!per_voice && !bank_mix && dev) ? out : b->d_out;
	const bool exch_db = bank_mix && mixdown && b->graph != KB_SY_SYNTHX;
	int exch_parity = 0;
	if (exch_db) {
		if (!b->d_out_alt) {
			KB_CUDA(cudaMalloc(&b->d_out_alt, (size_t)b->instances * b->channels * b->max_block * sizeof(float)));
			for (int k = 0; k < 2; k++) KB_CUDA(cudaEventCreateWithFlags(&b->ev_exch[k], cudaEventDisableTiming));
		}
		exch_parity = (int)(b->exch_count & 1);
		d_inst_dst = exch_parity ? b->d_out_alt : b->d_out;
		// the exchange kernel of block k - 2 read this buffer: waited for HERE, before the voice kernel (it finished a block ago), so that the
		// mix kernel stays the voice kernel's programmatic dependent.  A plain process() in between wrote d_out behind wait_prev_exchange().
		if (b->exch_count >= 2) KB_CUDA(cudaStreamWaitEvent(st, b->ev_exch[exch_parity], 0));
	}
	if (b->graph == KB_SY_SYNTHX) {
		const int pthreads = total * 132;
		kb_sx_prepare_kernel<<<(pthreads + 127) / 128, 128, 0, st>>>((KbSxVoice*)b->d_vstate, b->d_hdr, b->d_blk, b->voices, total, b->fs);
		kb_sx_adsr_kernel<<<(total + 31) / 32, 32, 0, st>>>((KbSxVoice*)b->d_vstate, b->d_hdr, b->d_adsr, n, total, b->fs);
		dim3 grid((n + 31) / 32, b->instances);
		b->prof_begin();
		kb_sx_render_kernel<<<grid, 32 * (KB_SX_PRODUCERS * KB_SX_WARPS_PER_PAIR + 1), sizeof(KbSxSmem), st>>>((const KbSxVoice*)b->d_vstate, b->d_hdr, b->d_adsr, per_voice ? d_voice_dst : d_inst_dst, n, b->voices, per_voice ? 1 : 0);
		b->prof_end();
		kb_sx_advance_kernel<<<(pthreads + 127) / 128, 128, 0, st>>>((KbSxVoice*)b->d_vstate, b->d_hdr, n, total);
		b->launches += 4;
	} else {
		b->prof_begin();
		// FM.k: the time-parallel kernel (kb_fm_tiled_kernel) is the default; KB_FM_TILED=0 keeps the lane-per-voice kernel for the whole
		// process (A/B measurement, same results — tools/fm_tiled_probe.py)
		static const bool fm_tiled = !getenv("KB_FM_TILED") || atoi(getenv("KB_FM_TILED")) != 0;
		if ((flags & KB_LANE_PER_VOICE) || (b->graph == KB_SY_FM && !fm_tiled) || b->graph >= KB_SY_BREAKPOINT) {   // (the Sine x envelope graphs: lane per voice)
			const int blocks = (total + 127) / 128;
			switch (b->graph) {
			case KB_SY_SUBTRACTIVE: case KB_SY_FILTER_K:
				kb_voice_kernel<KB_SY_SUBTRACTIVE, KbSubVoice><<<blocks, 128, 0, st>>>((KbSubVoice*)b->d_vstate, b->d_hdr, b->d_blk, d_voice_dst, n, b->voices, total, b->fs); break;
			case KB_SY_SUPERSAW:
				kb_voice_kernel<KB_SY_SUPERSAW, KbSsawVoice><<<blocks, 128, 0, st>>>((KbSsawVoice*)b->d_vstate, b->d_hdr, b->d_blk, d_voice_dst, n, b->voices, total, b->fs); break;
			case KB_SY_TB303:
				kb_voice_kernel<KB_SY_TB303, KbTbVoice><<<blocks, 128, 0, st>>>((KbTbVoice*)b->d_vstate, b->d_hdr, b->d_blk, d_voice_dst, n, b->voices, total, b->fs); break;
			case KB_SY_FM:
				kb_voice_kernel<KB_SY_FM, KbFmVoice><<<blocks, 128, 0, st>>>((KbFmVoice*)b->d_vstate, b->d_hdr, b->d_blk, d_voice_dst, n, b->voices, total, b->fs); break;
#define KB_LAUNCH_ESINE(VOICE)                                                                                                       \
	do {                                                                                                                             \
		static bool attr_set = false;                                                                                                \
		if (!attr_set) { cudaFuncSetAttribute(kb_esine_tiled_kernel<VOICE, 8, 544>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(KbEsSmem<VOICE, 8>)); attr_set = true; } \
		kb_esine_tiled_kernel<VOICE, 8, 544><<<(total + 7) / 8, 544, sizeof(KbEsSmem<VOICE, 8>), st>>>((VOICE*)b->d_vstate, b->d_hdr, b->d_blk, d_voice_dst, n, b->voices, total, b->fs); \
	} while (0)
			case KB_SY_BREAKPOINT: case KB_SY_RAMP: case KB_SY_RELEASE:      // time-parallel unless KB_LANE_PER_VOICE: one envelope lane per voice, thread per (voice, sample)
				if (flags & KB_LANE_PER_VOICE) kb_voice_kernel<KB_SY_BREAKPOINT, KbSenvVoice><<<blocks, 128, 0, st>>>((KbSenvVoice*)b->d_vstate, b->d_hdr, b->d_blk, d_voice_dst, n, b->voices, total, b->fs);
				else KB_LAUNCH_ESINE(KbSenvVoice);
				break;
			case KB_SY_AM:
				if (flags & KB_LANE_PER_VOICE) kb_voice_kernel<KB_SY_AM, KbSmodVoice><<<blocks, 128, 0, st>>>((KbSmodVoice*)b->d_vstate, b->d_hdr, b->d_blk, d_voice_dst, n, b->voices, total, b->fs);
				else KB_LAUNCH_ESINE(KbSmodVoice);
				break;
#undef KB_LAUNCH_ESINE
			case KB_SY_MOD_FM: case KB_SY_MOD_FM2:
				kb_voice_kernel<KB_SY_AM, KbSmodVoice><<<blocks, 128, 0, st>>>((KbSmodVoice*)b->d_vstate, b->d_hdr, b->d_blk, d_voice_dst, n, b->voices, total, b->fs); break;
			case KB_SY_ADDITIVE_SAW: case KB_SY_ADDITIVE_SQUARE: case KB_SY_ADDITIVE_NYQUIST:
				if (flags & KB_LANE_PER_VOICE) {
					kb_voice_kernel<KB_SY_ADDITIVE_SAW, KbAddVoice><<<blocks, 128, 0, st>>>((KbAddVoice*)b->d_vstate, b->d_hdr, b->d_blk, d_voice_dst, n, b->voices, total, b->fs);
				} else {                                   // time-parallel: thread = (voice, sample), then the phase advance
					dim3 grid((unsigned)std::max(1, std::min((n + 255) / 256, 16)), total);
					kb_additive_kernel<<<grid, 256, 0, st>>>((const KbAddVoice*)b->d_vstate, b->d_hdr, d_voice_dst, n, b->fs);
					kb_additive_advance_kernel<<<(total * 32 + 127) / 128, 128, 0, st>>>((KbAddVoice*)b->d_vstate, b->d_hdr, total, n, b->fs);
					b->launches++;
				}
				break;
			}
		} else {
			// voices per CTA: as many as still leave >= ~100 CTAs (the serial stages cost the same for any G), KB_TILE_G overrides
			int g = sub ? sub_g : (b->graph == KB_SY_SUPERSAW ? (total >= 1024 ? 8 : 2) : (total >= 800 ? 8 : 4));
			if (force_g) g = force_g;
#define KB_LAUNCH_TILED(KERNEL, SMEM, GG, NT, ...)                                                                                   \
	do {                                                                                                                             \
		static bool attr_set = false;                                                                                                \
		if (!attr_set) { cudaFuncSetAttribute(KERNEL<GG, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SMEM<GG>)); attr_set = true; } \
		KERNEL<GG, NT><<<(total + GG - 1) / GG, NT, sizeof(SMEM<GG>), st>>>(__VA_ARGS__);                                            \
	} while (0)
			if (sub) {
				KbSubVoice* vs = (KbSubVoice*)b->d_vstate;
				// KB_TILE_LAYOUT (A/B measurement, same results): 2 = default for 800..1599 voices, the filter warp alone on its SM
				// sub-partition and both envelopes in one warp (kb_tiled.cuh); 0 = serial roles spread over the sub-partitions
				// KB_C2_TRACE=<file> (measurement aid): per-role clock64() stamps of CTA 0 for every tick of the last launch
				static const char* c2_trace_path = getenv("KB_C2_TRACE");
				static long long* c2_trace = nullptr;
				if (c2_trace_path && !c2_trace) { KB_CUDA(cudaMalloc(&c2_trace, 12 * 64 * 2 * sizeof(long long))); }
				if (c2_trace) KB_CUDA(cudaMemsetAsync(c2_trace, 0, 12 * 64 * 2 * sizeof(long long), st));
#define KB_LAUNCH_SUB(GG, NT, LAY)                                                                                                    \
	do {                                                                                                                             \
		static bool attr_set = false;                                                                                                \
		if (!attr_set) { cudaFuncSetAttribute(kb_sub_tiled_kernel<GG, NT, LAY>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(KbSubSmem<GG>)); attr_set = true; } \
		kb_sub_tiled_kernel<GG, NT, LAY><<<(total + GG - 1) / GG, NT, sizeof(KbSubSmem<GG>), st>>>(vs, b->d_hdr, d_voice_dst, n, total, b->fs, c2_trace); \
	} while (0)
				// layout 3 (round 2): the same stages decoupled, kb_sub_flow_kernel; KB_TILE_G=7 -> 147 CTAs, KB_TILE_ASP0=1 -> envelope warp beside the filter warp
				static const int c2_variant = getenv("KB_C2_VARIANT") ? atoi(getenv("KB_C2_VARIANT")) : 0;
				static const int asp0 = getenv("KB_TILE_ASP0") ? atoi(getenv("KB_TILE_ASP0")) : 1;
#define KB_LAUNCH_FLOW(GG, ASP0)                                                                                                      \
	do {                                                                                                                             \
		static bool attr_set = false;                                                                                                \
		if (!attr_set) { cudaFuncSetAttribute(kb_sub_flow_kernel<GG, ASP0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(KbSubFlowSmem<GG>)); attr_set = true; } \
		kb_sub_flow_kernel<GG, ASP0><<<(total + GG - 1) / GG, 768, sizeof(KbSubFlowSmem<GG>), st>>>(vs, b->d_hdr, d_voice_dst, n, total, b->fs, b->staged, c2_trace, c2_variant); \
	} while (0)
#define KB_LAUNCH_MBAR(GG, CHK)                                                                                                       \
	do {                                                                                                                             \
		static bool attr_set = false;                                                                                                \
		if (!attr_set) { cudaFuncSetAttribute(kb_sub_mbar_kernel<GG, CHK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(KbSubMbarSmem<GG>)); attr_set = true; } \
		kb_sub_mbar_kernel<GG, CHK><<<(total + GG - 1) / GG, 768, sizeof(KbSubMbarSmem<GG>), st>>>(vs, b->d_hdr, d_voice_dst, n, total, b->fs, b->staged, c2_trace, c2_variant); \
	} while (0)
				if (sub_flow && layout >= 4) {
					if (c2_variant & 128) { if (g == 7) KB_LAUNCH_MBAR(7, true); else KB_LAUNCH_MBAR(8, true); }     // the self-checking instantiation
					else if (g == 7) KB_LAUNCH_MBAR(7, false); else KB_LAUNCH_MBAR(8, false);
					if (b->staged.count > 0) { KB_CUDA(cudaEventRecord(b->stage_done[b->staged_slot], st)); b->staged.count = 0; }
				}
				else if (sub_flow) {
					if (g == 7) { if (asp0) KB_LAUNCH_FLOW(7, true); else KB_LAUNCH_FLOW(7, false); }
					else { if (asp0) KB_LAUNCH_FLOW(8, true); else KB_LAUNCH_FLOW(8, false); }
					if (b->staged.count > 0) { KB_CUDA(cudaEventRecord(b->stage_done[b->staged_slot], st)); b->staged.count = 0; }
				}
				else if (g >= 16) KB_LAUNCH_SUB(16, 1024, 0);
				else if (g >= 7 && layout >= 2) KB_LAUNCH_SUB(8, 768, 2);
				else if (g >= 8) KB_LAUNCH_SUB(8, 512, 0);
				else if (g == 7) KB_LAUNCH_SUB(7, 544, 0);               // 448 worker threads cover a 7 x 128 tile in exactly two rounds
				else KB_LAUNCH_SUB(4, 320, 0);
#undef KB_LAUNCH_SUB
#undef KB_LAUNCH_FLOW
#undef KB_LAUNCH_MBAR
				if (c2_trace) {
					std::vector<long long> tr(12 * 64 * 2);
					KB_CUDA(cudaMemcpyAsync(tr.data(), c2_trace, tr.size() * sizeof(long long), cudaMemcpyDeviceToHost, st));
					KB_CUDA(cudaStreamSynchronize(st));
					if (FILE* f = fopen(c2_trace_path, "w")) {
						for (int r = 0; r < 12; r++) for (int k = 0; k < 64; k++) if (tr[(r * 64 + k) * 2] || tr[(r * 64 + k) * 2 + 1]) fprintf(f, "%d %d %lld %lld\n", r, k, tr[(r * 64 + k) * 2], tr[(r * 64 + k) * 2 + 1]);
						fclose(f);
					}
				}
			} else if (b->graph == KB_SY_SUPERSAW) {
				KbSsawVoice* vs = (KbSsawVoice*)b->d_vstate;
				if (g >= 8) KB_LAUNCH_TILED(kb_ssaw_tiled_kernel, KbSsawSmem, 8, 1024, vs, b->d_hdr, d_voice_dst, n, total, b->fs);
				else KB_LAUNCH_TILED(kb_ssaw_tiled_kernel, KbSsawSmem, 2, 1024, vs, b->d_hdr, d_voice_dst, n, total, b->fs);   // 992 workers: 7 x 2 x 128 items in two rounds
			} else if (b->graph == KB_SY_FM) {
				KB_LAUNCH_TILED(kb_fm_tiled_kernel, KbFmSmem, 8, 544, (KbFmVoice*)b->d_vstate, b->d_hdr, b->d_blk, d_voice_dst, n, b->voices, total, b->fs);   // 512 workers: an 8 x 128 tile in two rounds
			} else {
				KbTbVoice* vs = (KbTbVoice*)b->d_vstate;
				// (the ladder recurrence, one lane per voice and ~100 dependent cycles per sample, bounds this kernel for any G; giving it a
				// sub-partition of its own — layout 2 of the Subtractive kernel — measured no faster here)
				if (g >= 8) KB_LAUNCH_TILED(kb_tb_tiled_kernel, KbTbSmem, 8, 512, vs, b->d_hdr, b->d_blk, d_voice_dst, n, b->voices, total, b->fs);
				else KB_LAUNCH_TILED(kb_tb_tiled_kernel, KbTbSmem, 4, 320, vs, b->d_hdr, b->d_blk, d_voice_dst, n, b->voices, total, b->fs);
			}
#undef KB_LAUNCH_TILED
		}
		b->prof_end();
		b->launches++;
		if (!per_voice) {
			if (!exch_db) { rc = wait_prev_exchange(); if (rc) return rc; }
			// KB_MIX_FUSED=0 keeps the two-kernel mix (A/B measurement, same results)
			static const bool mix_fused = !getenv("KB_MIX_FUSED") || atoi(getenv("KB_MIX_FUSED")) != 0;
			if (mix_fused && C == 1 && (flags & KB_MIX_SUM)) {
				// one launch for the voice sums of every instance and — on one GPU — the bank mix (kb_mix_fused_kernel)
				static int smem_max = 0;
				if (!smem_max) { KB_CUDA(cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, b->device)); KB_CUDA(cudaFuncSetAttribute(kb_mix_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max)); }
				const size_t per_inst = (size_t)b->voices * KB_MIXF_TS * sizeof(float) + KB_MIXF_TS * sizeof(float) + (size_t)b->voices * sizeof(int);
				const int group = (int)std::max<size_t>(1, std::min<size_t>((size_t)b->instances, (size_t)smem_max / per_inst));
				const bool fuse_bank = bank_mix && !mixdown;
				// a page-locked, device-mapped host buffer takes the mix straight from the kernel's stores (16 KiB per block over PCIe inside the
				// kernel): no copy-engine operation queued behind the kernels (KB_ZERO_COPY_OUT=0: copy as before; same bytes)
				static const bool zero_copy = !getenv("KB_ZERO_COPY_OUT") || atoi(getenv("KB_ZERO_COPY_OUT")) != 0;
				if (fuse_bank) d_result_fused = dev ? out : (zero_copy && kb_host_ptr_mapped(out) ? out : b->d_mix);
				// programmatic dependent launch behind the voice kernel (KB_PDL=0: a plain launch; same results): the launch latency of this
				// kernel overlaps the voice kernel's tail
				static const bool pdl = !getenv("KB_PDL") || atoi(getenv("KB_PDL")) != 0;
				cudaLaunchConfig_t cfg = {};
				cfg.gridDim = dim3((unsigned)((n + KB_MIXF_TS - 1) / KB_MIXF_TS)); cfg.blockDim = dim3(1024); cfg.dynamicSmemBytes = group * per_inst; cfg.stream = st;
				cudaLaunchAttribute attr[1];
				attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; attr[0].val.programmaticStreamSerializationAllowed = 1;
				cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
				KB_CUDA(cudaLaunchKernelEx(&cfg, kb_mix_fused_kernel, (const float*)b->d_scratch, (const KbVoiceHdr*)b->d_hdr, d_inst_dst, d_result_fused, n, b->voices, b->instances, group));
			} else {
				dim3 grid((n + 255) / 256, b->instances);
				kb_mix_kernel<<<grid, 256, 0, st>>>(b->d_scratch, b->d_hdr, d_inst_dst, n, b->voices, (flags & KB_MIX_SUM) ? 1 : 0);
			}
			b->launches++;
		}
	}
	float* d_result = per_voice ? d_voice_dst : d_inst_dst;
	if (bank_mix && mixdown) {
		// the bank mix goes straight into rank 0's arena (kb_mixdown_step_kernel); nothing is returned through `out`
		if (!b->mix_stream) { KB_CUDA(cudaStreamCreateWithFlags(&b->mix_stream, cudaStreamNonBlocking)); KB_CUDA(cudaEventCreateWithFlags(&b->ev_mix_in, cudaEventDisableTiming)); }
		KB_CUDA(cudaEventRecord(b->ev_mix_in, st));
		KB_CUDA(cudaStreamWaitEvent(b->mix_stream, b->ev_mix_in, 0));
		rc = mixdown_step(mixdown, d_inst_dst, b->instances, (size_t)C * n, C * n, out_prev, b->mix_stream); if (rc) return rc;
		if (exch_db) { KB_CUDA(cudaEventRecord(b->ev_exch[exch_parity], b->mix_stream)); b->exch_count++; }
		b->last_mixdown = mixdown;
		b->launches++;
		b->host_stale = true; b->hdr_stale = true; b->vstate_stale = true;
		return KB_OK;
	}
	if (bank_mix && d_result_fused) d_result = d_result_fused;
	else if (bank_mix) {
		d_result = dev ? out : b->d_mix;
		kb_bank_mix_kernel<<<(C * n + 255) / 256, 256, 0, st>>>(b->d_out, d_result, C * n, b->instances);
		b->launches++;
	}
	KB_CUDA(cudaGetLastError());
	b->host_stale = true; b->hdr_stale = true; b->vstate_stale = true;
	if (!dev) {
		if (d_result != out) KB_CUDA(cudaMemcpyAsync(out, d_result, out_floats * sizeof(float), cudaMemcpyDeviceToHost, st));
		if (!(flags & KB_ASYNC_HOST)) KB_CUDA(cudaStreamSynchronize(st));
		b->d2h_bytes += (long long)(out_floats * sizeof(float));
	}
	return KB_OK;
}

// ==================================================================================== multi-GPU mix-down
// (include/klang_b200.h: kb_mixdown_*)  Arena on rank 0: float slots[2][world][max_floats]; then unsigned flags[world] (the
// last step each rank has published) and unsigned consumed (the last step rank 0 has summed).  All waits are device-side
// spins on system-scope volatile words, reached over NVLink by the peers.
__global__ void kb_mixdown_wait_free_kernel(volatile unsigned* consumed, unsigned need) {
	while ((int)(*consumed - need) < 0) __nanosleep(200);
}
__global__ void kb_mixdown_copy_kernel(const float* __restrict__ src, float* __restrict__ dst, int count) {
	for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x) dst[i] = src[i];
}
__global__ void kb_mixdown_publish_kernel(volatile unsigned* flag, unsigned step) {
	__threadfence_system();
	*flag = step;
}
__global__ void __launch_bounds__(1024) kb_mixdown_collect_kernel(const float* __restrict__ slots, size_t slot_floats, volatile unsigned* flags, volatile unsigned* consumed,
                                                                 unsigned step, int world, float* __restrict__ dst, int count) {
	if ((int)threadIdx.x < world) while ((int)(flags[threadIdx.x] - step) < 0) __nanosleep(100);
	__syncthreads();
	__threadfence_system();
	for (int i = threadIdx.x; i < count; i += blockDim.x) {
		float acc = __ldcv(slots + i);                                    // (peer-written memory: never from a stale cache line)
		for (int r = 1; r < world; r++) acc += __ldcv(slots + (size_t)r * slot_floats + i);
		dst[i] = acc;
	}
	__syncthreads();
	if (threadIdx.x == 0) { __threadfence_system(); *consumed = step; }
}
extern "C" kb_mixdown* kb_mixdown_create(int device, int world, int rank, int max_floats) {
	if (world < 1 || world > 1024 || rank < 0 || rank >= world || max_floats < 1) { kb_fail(KB_EINVAL, "kb_mixdown_create: bad argument"); return nullptr; }
	if (kb_device_count() < 1) { kb_fail(KB_ENODEV, "kb_mixdown_create: no CUDA device"); return nullptr; }
	if (cudaSetDevice(device) != cudaSuccess) { kb_fail(KB_ECUDA, "kb_mixdown_create: cudaSetDevice failed"); return nullptr; }
	kb_mixdown* m = new kb_mixdown();
	m->device = device; m->world = world; m->rank = rank; m->max_floats = max_floats;
	if (rank == 0) {
		if (cudaMalloc(&m->arena, m->arena_bytes()) != cudaSuccess || cudaMemset(m->arena, 0, m->arena_bytes()) != cudaSuccess) {
			kb_fail(KB_ECUDA, "kb_mixdown_create: arena allocation failed"); delete m; return nullptr;
		}
	}
	return m;
}
extern "C" void kb_mixdown_destroy(kb_mixdown* m) {
	if (!m) return;
	cudaSetDevice(m->device);
	cudaDeviceSynchronize();
	if (m->arena) { if (m->mapped) cudaIpcCloseMemHandle(m->arena); else cudaFree(m->arena); }
	cudaFree(m->d_tickets);
	if (m->ev_done) cudaEventDestroy(m->ev_done);
	for (int k = 0; k < 4; k++) if (m->ev_ring[k]) cudaEventDestroy(m->ev_ring[k]);
	delete m;
}
extern "C" int kb_mixdown_export(kb_mixdown* m, void* handle) {
	if (!m || !handle || m->rank != 0) return kb_fail(KB_EINVAL, "kb_mixdown_export: rank 0 only");
	static_assert(sizeof(cudaIpcMemHandle_t) == KB_IPC_HANDLE_BYTES, "IPC handle size");
	KB_CUDA(cudaSetDevice(m->device));
	KB_CUDA(cudaIpcGetMemHandle((cudaIpcMemHandle_t*)handle, m->arena));
	return KB_OK;
}
extern "C" int kb_mixdown_import(kb_mixdown* m, const void* handle) {
	if (!m || !handle || m->rank == 0 || m->arena) return kb_fail(KB_EINVAL, "kb_mixdown_import: ranks > 0, once");
	KB_CUDA(cudaSetDevice(m->device));
	cudaIpcMemHandle_t h; memcpy(&h, handle, sizeof(h));
	void* p = nullptr;
	KB_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
	m->arena = (unsigned char*)p; m->mapped = true;
	return KB_OK;
}
extern "C" float* kb_mixdown_acquire(kb_mixdown* m, void* stream) {
	if (!m || !m->arena) { kb_fail(KB_EINVAL, "kb_mixdown_acquire: arena not mapped"); return nullptr; }
	if (cudaSetDevice(m->device) != cudaSuccess) { kb_fail(KB_ECUDA, "kb_mixdown_acquire: cudaSetDevice failed"); return nullptr; }
	const unsigned step = ++m->step;
	// the slot of this parity was last used by step - 2: rank 0 must have summed that step
	if (step > 2) kb_mixdown_wait_free_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(m->consumed(), step - 2);
	return m->slot(step, m->rank);
}
extern "C" int kb_mixdown_publish(kb_mixdown* m, void* stream) {
	if (!m || !m->arena || m->step == 0) return kb_fail(KB_EINVAL, "kb_mixdown_publish: nothing acquired");
	KB_CUDA(cudaSetDevice(m->device));
	kb_mixdown_publish_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(m->flags() + m->rank, m->step);
	KB_CUDA(cudaGetLastError());
	return KB_OK;
}
extern "C" int kb_mixdown_put(kb_mixdown* m, const float* src, int count, void* stream) {
	if (!m || !src || count < 0 || count > m->max_floats) return kb_fail(KB_EINVAL, "kb_mixdown_put: bad argument");
	float* slot = kb_mixdown_acquire(m, stream);
	if (!slot) return KB_EINVAL;
	// a kernel storing through the peer mapping (cudaMemcpyAsync into IPC-mapped memory measured ~100 us per call)
	if (count > 0) kb_mixdown_copy_kernel<<<(count + 1023) / 1024, 256, 0, (cudaStream_t)stream>>>(src, slot, count);
	return kb_mixdown_publish(m, stream);
}
extern "C" int kb_mixdown_collect(kb_mixdown* m, float* dst, int count, void* stream) {
	if (!m || !dst || m->rank != 0 || count < 0 || count > m->max_floats || m->step == 0) return kb_fail(KB_EINVAL, "kb_mixdown_collect: bad argument (rank 0 only, count <= max_floats)");
	KB_CUDA(cudaSetDevice(m->device));
	// (a fused step may still be running on another stream and writes the caller's out_prev buffer: order this collect behind it)
	if (m->step_pending) { KB_CUDA(cudaStreamWaitEvent((cudaStream_t)stream, m->ev_done, 0)); m->step_pending = false; }
	kb_mixdown_collect_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(m->slot(m->step, 0), (size_t)m->max_floats, m->flags(), m->consumed(), m->step, m->world, dst, count);
	m->collected = m->step;
	KB_CUDA(cudaGetLastError());
	return KB_OK;
}

// The fused step (kb_mixdown_step / kb_synth_bank_process_mixdown): ONE kernel per block and rank, no separate wait / publish / collect
// launches and no collective library on the data path.
//   (a) thread 0 of every CTA waits until rank 0 has consumed the step that last used this slot parity (back-pressure; ranks > 0 poll
//       the word over NVLink only when they run two blocks ahead);
//   (b) the rank's contribution — the in-order fp32 sum of `rows` rows of `src` (the per-instance Synth outputs: the bank mix,
//       klang.h:4842-4848 applied across instances) — is stored into the rank's slot, peer memory on ranks > 0;
//   (c) the last CTA to finish raises the rank's flag behind a system-scope fence;
//   (d) on rank 0 the same kernel then sums the slots of the PREVIOUS step in rank order into out_prev and marks that step consumed:
//       the exchange of block k overlaps the voice kernels of block k + 1, and a late rank never stalls the others' current block.
__global__ void __launch_bounds__(256, 8) kb_mixdown_step_kernel(const float* __restrict__ src, int rows, size_t row_stride, int count, float* __restrict__ slot,
                                                              volatile unsigned* flags, volatile unsigned* consumed, unsigned* tickets, unsigned step, int rank, int world,
                                                              const float* prev_slots, size_t slot_floats, float* __restrict__ out_prev, int do_prev) {
	__shared__ bool s_last;
	if (threadIdx.x == 0 && step > 2) while ((int)(*consumed - (step - 2)) < 0) __nanosleep(200);
	__syncthreads();
	// (128-bit accesses with the rows' loads in flight together where the layout allows: one small CTA moves the whole mix)
	const bool vec = (count & 3) == 0 && (row_stride & 3) == 0 && ((reinterpret_cast<size_t>(src) | reinterpret_cast<size_t>(slot)) & 15) == 0;
	if (vec) {
		const float4* s4 = reinterpret_cast<const float4*>(src);
		float4* d4 = reinterpret_cast<float4*>(slot);
		const size_t rs4 = row_stride / 4;
		for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < count / 4; i += gridDim.x * blockDim.x) {
			float4 acc = s4[i];
			#pragma unroll 4
			for (int k = 1; k < rows; k++) { const float4 v = s4[(size_t)k * rs4 + i]; acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w; }   // row order kept per sample
			d4[i] = acc;
		}
	} else
	for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x) {
		float acc = src[i];
		for (int k = 1; k < rows; k++) acc += src[(size_t)k * row_stride + i];
		slot[i] = acc;
	}
	__threadfence_system();
	__syncthreads();
	if (threadIdx.x == 0) {
		s_last = atomicAdd(&tickets[0], 1u) == gridDim.x - 1;
		if (s_last) { tickets[0] = 0u; __threadfence_system(); flags[rank] = step; }
	}
	if (!do_prev) return;
	for (int r = threadIdx.x; r < world; r += blockDim.x) while ((int)(flags[r] - (step - 1)) < 0) __nanosleep(100);
	__syncthreads();
	__threadfence_system();
	const bool vec_prev = (count & 3) == 0 && (slot_floats & 3) == 0 && ((reinterpret_cast<size_t>(prev_slots) | reinterpret_cast<size_t>(out_prev)) & 15) == 0;
	if (vec_prev) {
		const float4* p4 = reinterpret_cast<const float4*>(prev_slots);
		const size_t sf4 = slot_floats / 4;
		for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < count / 4; i += gridDim.x * blockDim.x) {
			float4 acc = __ldcv(p4 + i);                                   // (peer-written memory: never from a stale cache line)
			#pragma unroll 4
			for (int r = 1; r < world; r++) { const float4 v = __ldcv(p4 + (size_t)r * sf4 + i); acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w; }   // rank order
			reinterpret_cast<float4*>(out_prev)[i] = acc;
		}
	} else
	for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x) {
		float acc = __ldcv(prev_slots + i);                                // (peer-written memory: never from a stale cache line)
		for (int r = 1; r < world; r++) acc += __ldcv(prev_slots + (size_t)r * slot_floats + i);
		out_prev[i] = acc;
	}
	__syncthreads();
	if (threadIdx.x == 0) {
		__threadfence_system();
		if (atomicAdd(&tickets[1], 1u) == gridDim.x - 1) { tickets[1] = 0u; __threadfence_system(); *consumed = step - 1; }
	}
}
static int mixdown_step(kb_mixdown* m, const float* src, int rows, size_t row_stride, int count, float* out_prev, cudaStream_t stream) {
	if (!m->d_tickets) { KB_CUDA(cudaMalloc(&m->d_tickets, 2 * sizeof(unsigned))); KB_CUDA(cudaMemset(m->d_tickets, 0, 2 * sizeof(unsigned))); }
	const unsigned step = ++m->step;
	const int do_prev = (m->rank == 0 && step > 1 && m->collected < step - 1 && out_prev) ? 1 : 0;
	// ONE CTA of 128 threads and at most 32 registers by default (KB_MIXDOWN_CTAS / KB_MIXDOWN_THREADS: A/B).  The exchange kernel runs beside
	// the next block's voice kernel, whose CTAs take 61440 of an SM's 65536 registers: this CTA fits into the 4096 that are left on ANY SM, so it
	// never keeps a voice CTA out while it waits for the peers' flags (32 CTAs of 256 threads did, on 32 SMs: +10 us per block at N > 1)
	static const int exch_ctas = getenv("KB_MIXDOWN_CTAS") ? std::max(1, atoi(getenv("KB_MIXDOWN_CTAS"))) : 1;
	static const int exch_threads = getenv("KB_MIXDOWN_THREADS") ? std::min(256, std::max(32, atoi(getenv("KB_MIXDOWN_THREADS")))) : 128;
	kb_mixdown_step_kernel<<<std::max(1, std::min(exch_ctas, (count + 255) / 256)), exch_threads, 0, stream>>>(src, rows, row_stride, count, m->slot(step, m->rank), m->flags(), m->consumed(), m->d_tickets,
	                                                                                          step, m->rank, m->world, m->slot(step - 1, 0), (size_t)m->max_floats, out_prev, do_prev);
	if (do_prev) m->collected = step - 1;
	KB_CUDA(cudaGetLastError());
	if (!m->ev_done) KB_CUDA(cudaEventCreateWithFlags(&m->ev_done, cudaEventDisableTiming));
	KB_CUDA(cudaEventRecord(m->ev_done, stream));
	if (!m->ev_ring[step & 3]) KB_CUDA(cudaEventCreateWithFlags(&m->ev_ring[step & 3], cudaEventDisableTiming));
	KB_CUDA(cudaEventRecord(m->ev_ring[step & 3], stream));
	m->step_pending = true;
	return KB_OK;
}
// the host waits until the exchange kernel of the fused step `back` steps before the last one has finished (back = 0 .. 3): what a host that
// rotates its out_prev buffers does before it hands one out again, without joining the steps still in flight
extern "C" int kb_mixdown_host_wait(kb_mixdown* m, int back) {
	if (!m || back < 0 || back > 3) return kb_fail(KB_EINVAL, "kb_mixdown_host_wait: bad argument (back = 0 .. 3)");
	if ((unsigned)back >= m->step) return KB_OK;                         // that step never ran
	cudaEvent_t ev = m->ev_ring[(m->step - (unsigned)back) & 3];
	if (ev) { KB_CUDA(cudaSetDevice(m->device)); KB_CUDA(cudaEventSynchronize(ev)); }
	return KB_OK;
}
extern "C" int kb_synth_bank_process_mixdown(kb_synth_bank* b, kb_mixdown* m, float* out_prev, int n, unsigned flags) {
	if (!b || !m || !m->arena || n < 1 || n > b->max_block || b->channels * n > m->max_floats || (flags & KB_PER_VOICE))
		return kb_fail(KB_EINVAL, "kb_synth_bank_process_mixdown: bad argument (arena mapped? channels * n <= max_floats?)");
	if (m->rank == 0 && m->step > m->collected + 1) return kb_fail(KB_EINVAL, "kb_synth_bank_process_mixdown: rank 0 must pass out_prev on every step (or collect) so the slots are consumed");
	if (m->device != b->device) return kb_fail(KB_EINVAL, "kb_synth_bank_process_mixdown: bank and mix-down live on different devices");
	return sy_process(b, (float*)b->d_mix, n, flags | KB_BANK_MIX | KB_DEVICE_PTR, m, out_prev);
}
// the block's events, then kb_synth_bank_process_mixdown: one call per block for a host that drives a sharded bank
extern "C" int kb_synth_bank_step_mixdown(kb_synth_bank* b, int count, const kb_note_event* events, kb_mixdown* m, float* out_prev, int n, unsigned flags) {
	int rc = kb_synth_bank_events(b, count, events); if (rc) return rc;
	return kb_synth_bank_process_mixdown(b, m, out_prev, n, flags);
}
// make `stream` wait for the exchange kernel of the last fused step (it runs on the bank's side stream): what a consumer of `out_prev` queues
// before it reads the buffer, without tying the bank's own stream to the exchange
extern "C" int kb_mixdown_stream_wait(kb_mixdown* m, void* stream) {
	if (!m) return kb_fail(KB_EINVAL, "kb_mixdown_stream_wait: null mix-down");
	KB_CUDA(cudaSetDevice(m->device));
	if (m->step_pending && m->ev_done) KB_CUDA(cudaStreamWaitEvent((cudaStream_t)stream, m->ev_done, 0));
	return KB_OK;
}
extern "C" int kb_mixdown_step(kb_mixdown* m, const float* src, int count, float* out_prev, void* stream) {
	if (!m || !m->arena || !src || count < 1 || count > m->max_floats) return kb_fail(KB_EINVAL, "kb_mixdown_step: bad argument (arena mapped? count <= max_floats?)");
	if (m->rank == 0 && m->step > m->collected + 1) return kb_fail(KB_EINVAL, "kb_mixdown_step: rank 0 must pass out_prev on every step (or collect) so the slots are consumed");
	KB_CUDA(cudaSetDevice(m->device));
	return mixdown_step(m, src, 1, 0, count, out_prev, (cudaStream_t)stream);
}

// ============================================================================================ primitives
struct DevBuf {
	void* p = nullptr;
	DevBuf(size_t bytes, const void* src = nullptr) { if (cudaMalloc(&p, bytes ? bytes : 1) != cudaSuccess) p = nullptr; else if (src) cudaMemcpy(p, src, bytes, cudaMemcpyHostToDevice); }
	~DevBuf() { cudaFree(p); }
	template <class T> T* as() { return (T*)p; }
};
static int prim_finish(const char* what) {
	cudaError_t e = cudaDeviceSynchronize();
	if (e == cudaSuccess) e = cudaGetLastError();
	if (e != cudaSuccess) return kb_fail(KB_ECUDA, std::string(what) + ": " + cudaGetErrorString(e));
	return KB_OK;
}
// Wavetable::operator=(Oscillator) fills the table with a Basic osc at fs/size Hz on the host (klang.h:3645-3650); kind 10 Sine, 11 Saw
static void prim_fill_wavetable(int kind, const KbFs& F, float* table) {
	KbBasicOsc o; kb_bosc_init(o); kb_bosc_set_f(F, o, F.f / 2048);
	for (int s = 0; s < 2048; s++) {
		table[s] = (kind == 10) ? ::sinf(o.position + o.offset) : (o.position * KB_PI_INV_F - 1.f);
		kb_bosc_advance(o);
	}
}
extern "C" int kb_prim_wavetable(int kind, float fs, float* out) {
	if (kb_device_count() < 1) return kb_fail(KB_ENODEV, "no CUDA device");
	if ((kind != 10 && kind != 11) || !out || !(fs > 0)) return kb_fail(KB_EINVAL, "kb_prim_wavetable: kind 10 (Wavetables::Sine) or 11 (Wavetables::Saw)");
	std::vector<float> table(2048);
	prim_fill_wavetable(kind, kb_make_fs(fs), table.data());
	DevBuf dtab(sizeof(float) * 2048, table.data()), dout(sizeof(float) * 2048);
	kb_prim_wavetable_kernel<<<8, 256>>>(dtab.as<float>(), dout.as<float>());
	int rc = prim_finish("kb_prim_wavetable"); if (rc) return rc;
	cudaMemcpy(out, dout.p, sizeof(float) * 2048, cudaMemcpyDeviceToHost);
	return KB_OK;
}
extern "C" int kb_prim_stereo_delay(int n, const float* inl, const float* inr, const float* df, float* outl, float* outr) {
	if (kb_device_count() < 1) return kb_fail(KB_ENODEV, "no CUDA device");
	if (n < 0 || !inl || !inr || !df || !outl || !outr) return kb_fail(KB_EINVAL, "kb_prim_stereo_delay: bad argument");
	for (int s = 0; s < n; s++) if (!(df[s] >= 0.f && df[s] < 999.f)) return kb_fail(KB_EINVAL, "kb_prim_stereo_delay: delay outside the 1000-sample line");
	if (n == 0) return KB_OK;
	const size_t B = sizeof(float) * n;
	DevBuf dl(B, inl), dr(B, inr), ddf(B, df), rl(sizeof(float) * 1001), rr(sizeof(float) * 1001), ol(B), orr(B);
	kb_prim_stereo_delay_kernel<<<1, 32>>>(n, dl.as<float>(), dr.as<float>(), ddf.as<float>(), rl.as<float>(), rr.as<float>(), ol.as<float>(), orr.as<float>());
	int rc = prim_finish("kb_prim_stereo_delay"); if (rc) return rc;
	cudaMemcpy(outl, ol.p, B, cudaMemcpyDeviceToHost); cudaMemcpy(outr, orr.p, B, cudaMemcpyDeviceToHost);
	return KB_OK;
}
extern "C" int kb_prim_control_smooth(float lo, float hi, float initial, int n, const float* values, float* out) {
	if (kb_device_count() < 1) return kb_fail(KB_ENODEV, "no CUDA device");
	if (n < 0 || !values || !out) return kb_fail(KB_EINVAL, "kb_prim_control_smooth: bad argument");
	if (n == 0) return KB_OK;
	DevBuf dv(sizeof(float) * n, values), dout(sizeof(float) * n);
	kb_prim_control_smooth_kernel<<<1, 32>>>(lo, hi, initial, n, dv.as<float>(), dout.as<float>());
	int rc = prim_finish("kb_prim_control_smooth"); if (rc) return rc;
	cudaMemcpy(out, dout.p, sizeof(float) * n, cudaMemcpyDeviceToHost);
	return KB_OK;
}
extern "C" int kb_prim_envelope_at(int npts, const float* xy, int n, const float* t, float* out) {
	if (kb_device_count() < 1) return kb_fail(KB_ENODEV, "no CUDA device");
	if (npts < 0 || npts > KB_ENV_MAXPTS || !xy || n < 0 || !t || !out) return kb_fail(KB_EINVAL, "kb_prim_envelope_at: bad argument");
	if (n == 0) return KB_OK;
	DevBuf dxy(sizeof(float) * 2 * (npts ? npts : 1), xy), dt(sizeof(float) * n, t), dout(sizeof(float) * n);
	kb_prim_envelope_at_kernel<<<std::min(148, (n + 127) / 128), 128>>>(npts, dxy.as<float>(), n, dt.as<float>(), dout.as<float>());
	int rc = prim_finish("kb_prim_envelope_at"); if (rc) return rc;
	cudaMemcpy(out, dout.p, sizeof(float) * n, cudaMemcpyDeviceToHost);
	return KB_OK;
}
extern "C" int kb_prim_osc(int kind, int nargs, float f, float phase, float duty, float fs, int n, float* out) {
	if (kb_device_count() < 1) return kb_fail(KB_ENODEV, "no CUDA device");
	if (!out || n < 0 || kind < 0 || kind > 13) return kb_fail(KB_EINVAL, "kb_prim_osc: unsupported kind");
	const KbFs F = kb_make_fs(fs);
	if (kind >= 12) {   // Basic::Noise / Fast::Noise: the device continues the process's libc rand() stream (kb_rand.h)
		KbRand g;
		if (!kb_rand_capture(g)) return kb_fail(KB_EINVAL, "kb_prim_osc: libc rand() is not running its default (TYPE_3) generator");
		DevBuf dn(sizeof(float) * (n ? n : 1));
		kb_prim_noise_kernel<<<1, 32>>>(kind == 13 ? 1 : 0, g, n, dn.as<float>());
		int nrc = prim_finish("kb_prim_osc"); if (nrc) return nrc;
		cudaMemcpy(out, dn.p, sizeof(float) * n, cudaMemcpyDeviceToHost);
		kb_rand_jump(g, (unsigned long long)n);          // the draws the device consumed
		kb_rand_commit(g);
		return KB_OK;
	}
	std::vector<float> table(2048, 0.f);
	if (kind >= 10) prim_fill_wavetable(kind, F, table.data());
	DevBuf dout(sizeof(float) * n), dtab(sizeof(float) * 2048, table.data());
	kb_prim_osc_kernel<<<1, 32>>>(kind, nargs, f, phase, duty, F, n, dout.as<float>(), dtab.as<float>());
	int rc = prim_finish("kb_prim_osc"); if (rc) return rc;
	cudaMemcpy(out, dout.p, sizeof(float) * n, cudaMemcpyDeviceToHost);
	return KB_OK;
}
// klang::Sample on the device over a caller's table (host memory; typically what kb_wav_decode produced)
extern "C" int kb_prim_sample(const float* table, int size, int nargs, float f, float phase, int n, float* out) {
	(void)f;                                                   // (Sample::set keeps the frequency but plays at one sample per tick, klang.h:3697-3700)
	if (kb_device_count() < 1) return kb_fail(KB_ENODEV, "no CUDA device");
	if (!table || !out || size < 1 || n < 0 || nargs < 1 || nargs > 2) return kb_fail(KB_EINVAL, "kb_prim_sample: bad argument");
	if (nargs == 2 && !(phase >= 0.f && phase * (float)44100 + 1.f <= (float)size)) return kb_fail(KB_EINVAL, "kb_prim_sample: the start phase lies outside the table");
	if (n == 0) return KB_OK;
	std::vector<float> padded((size_t)size + 2, 0.f);
	memcpy(padded.data(), table, sizeof(float) * (size_t)size);
	DevBuf dtab(sizeof(float) * padded.size(), padded.data()), dout(sizeof(float) * n);
	kb_prim_sample_kernel<<<1, 32>>>(dtab.as<float>(), size, nargs, phase, n, dout.as<float>());
	int rc = prim_finish("kb_prim_sample"); if (rc) return rc;
	cudaMemcpy(out, dout.p, sizeof(float) * n, cudaMemcpyDeviceToHost);
	return KB_OK;
}
// File::WAV::load + operator>> (klang.h:5997-6085) — host code in the reference too: the decoded floats are what a Sample (or any table) is fed.
// RIFF/WAVE header, chunk walk with the sizes as written (no pad byte), data->size / BlockAlign samples from CONSECUTIVE elements (a stereo file
// yields its first half, interleaved — as there): 8-bit unsigned (x - 128) / 255, 16- and 32-bit signed x / 2^15, x / 2^31, 32-bit float as is;
// any other encoding leaves zeros.  Returns the sample count (copies min(count, max_samples)); info = { NumChannels, SampleRate, BitsPerSample }.
extern "C" int kb_wav_decode(const void* image, long long nbytes, float* out, int max_samples, int* info) {
	const unsigned char* bytes = (const unsigned char*)image;
	if (!bytes || (!out && max_samples > 0) || max_samples < 0) return kb_fail(KB_EINVAL, "kb_wav_decode: bad argument");
	if (nbytes < 12 || memcmp(bytes, "RIFF", 4) || memcmp(bytes + 8, "WAVE", 4)) return kb_fail(KB_EINVAL, "kb_wav_decode: not a RIFF/WAVE image");
	long long at = 12, fmt = -1, data = -1;
	while (at + 8 <= nbytes) {
		uint32_t size; memcpy(&size, bytes + at + 4, 4);
		if (!memcmp(bytes + at, "fmt ", 4)) fmt = at;
		else if (!memcmp(bytes + at, "data", 4)) { if (at + (long long)size > nbytes) return kb_fail(KB_EINVAL, "kb_wav_decode: corrupt data chunk"); data = at; }
		at += 8 + (long long)size;
	}
	if (fmt < 0 || data < 0 || fmt + 24 > nbytes) return kb_fail(KB_EINVAL, "kb_wav_decode: fmt or data chunk missing");
	uint16_t af, ch, align, bits; uint32_t rate, dsize;
	memcpy(&af, bytes + fmt + 8, 2); memcpy(&ch, bytes + fmt + 10, 2); memcpy(&rate, bytes + fmt + 12, 4);
	memcpy(&align, bytes + fmt + 20, 2); memcpy(&bits, bytes + fmt + 22, 2); memcpy(&dsize, bytes + data + 4, 4);
	if (!align) return kb_fail(KB_EINVAL, "kb_wav_decode: BlockAlign is zero");
	const long long count = dsize / align;
	if (count > 0x7fffffff) return kb_fail(KB_EINVAL, "kb_wav_decode: too long");
	if (info) { info[0] = ch; info[1] = (int)rate; info[2] = bits; }
	const bool pcm = af == 1 && (bits == 8 || bits == 16 || bits == 32), flt = af == 3 && bits == 32;
	if ((pcm || flt) && data + 8 + count * (bits / 8) > nbytes) return kb_fail(KB_EINVAL, "kb_wav_decode: data chunk shorter than its size");   // (the reference would read past the image)
	const unsigned char* d = bytes + data + 8;
	for (long long i = 0; i < count && i < max_samples; i++) {
		if (pcm && bits == 8) out[i] = ((float)d[i] - 128u) * (1.f / 255);
		else if (pcm && bits == 16) { int16_t v; memcpy(&v, d + 2 * i, 2); out[i] = (float)v * (1.f / 32768u); }
		else if (pcm) { int32_t v; memcpy(&v, d + 4 * i, 4); out[i] = (float)v * (1.f / 2147483648u); }
		else if (flt) memcpy(out + i, d + 4 * i, 4);
		else out[i] = 0.f;
	}
	return (int)count;
}
extern "C" int kb_prim_delay(int n, const float* in, const int* di, const float* df, const float* set_at,
                             float* out_i, float* out_f, float* out_p, float* out_l) {
	if (kb_device_count() < 1) return kb_fail(KB_ENODEV, "no CUDA device");
	if (n < 0 || !in || !di || !df || !set_at || !out_i || !out_f || !out_p || !out_l) return kb_fail(KB_EINVAL, "kb_prim_delay: bad argument");
	for (int s = 0; s < n; s++)
		if (di[s] < 0 || di[s] >= 1000 || !(df[s] >= 0.f && df[s] < 999.f)) return kb_fail(KB_EINVAL, "kb_prim_delay: delay outside the 1000-sample line");
	if (n == 0) return KB_OK;
	const size_t B = sizeof(float) * n;
	DevBuf din(B, in), ddi(sizeof(int) * n, di), ddf(B, df), dset(B, set_at), dring(sizeof(float) * 1001), oi(B), of(B), op(B), ol(B);
	kb_prim_delay_kernel<<<1, 32>>>(n, din.as<float>(), ddi.as<int>(), ddf.as<float>(), dset.as<float>(), dring.as<float>(),
	                                oi.as<float>(), of.as<float>(), op.as<float>(), ol.as<float>());
	int rc = prim_finish("kb_prim_delay"); if (rc) return rc;
	cudaMemcpy(out_i, oi.p, B, cudaMemcpyDeviceToHost); cudaMemcpy(out_f, of.p, B, cudaMemcpyDeviceToHost);
	cudaMemcpy(out_p, op.p, B, cudaMemcpyDeviceToHost); cudaMemcpy(out_l, ol.p, B, cudaMemcpyDeviceToHost);
	return KB_OK;
}
extern "C" int kb_prim_filter(int kind, int nset, const float* f, const float* Q, float fs, int n, const float* in, float* out, float* coeffs) {
	if (kb_device_count() < 1) return kb_fail(KB_ENODEV, "no CUDA device");
	const bool onepole = kind == 2 || kind == 3 || kind == 7, host_set = kind >= 12;     // kinds whose set() runs on the host (libm)
	if (kind < 0 || kind > 16 || !f || !in || !out || !coeffs || nset < 0 || nset > n || (host_set && nset > 1) || (kind >= 11 && !Q))
		return kb_fail(KB_EINVAL, "kb_prim_filter: unsupported");
	const KbFs F = kb_make_fs(fs);
	float4 hc = make_float4(0.f, 0.f, 0.05f, 0.f);
	if (kind == 12 && nset == 1) {               // Modal::set(f, decay)  klang.h:5832-5845 (exp / cos of floats: expf / cosf)
		const float w = f[0] * F.w;
		const float d = kb_clampf(::expf(-KB_PI_F / (Q[0] * F.f)), 1e-6f, 0.9999f);
		hc.x = 2.f * d * kb_clampf(::cosf(w), -0.9999f, 0.9999f);
		hc.y = -d * d;
	} else if (kind >= 13 && kind <= 16) {       // Follower() / Window() { set(0.01f, 0.1f); } then AR::set(attack, release)  klang.h:5871-5878, 5882-5885, 5912-5918
		float attack = 0.01f, release = 0.1f;
		if (nset == 1) { attack = f[0]; release = Q[0]; }
		hc.x = 1.f - (attack == 0.f ? 0.f : ::expf(-1.0f / (F.f * attack)));
		hc.y = 1.f - (release == 0.f ? 0.f : ::expf(-1.0f / (F.f * release)));
	}
	KbOnePole op; kb_onepole_construct(op, kind == 2 ? KB_OP_LPF : kind == 3 ? KB_OP_HPF : KB_OP_BW1);
	if (onepole && nset == 1) kb_onepole_set(F, op, f[0]);
	std::vector<float> op_sets;                   // set() every sample: OnePole / Butterworth<1> coefficients are host libm code (klang.h:5508-5512, 5786-5793)
	if (onepole && nset > 1) {
		KbOnePole h = op;
		for (int s = 0; s < nset; s++) { kb_onepole_set(F, h, f[s]); op_sets.push_back(h.b0); op_sets.push_back(h.b1); op_sets.push_back(h.a1); }
	}
	DevBuf dsets(sizeof(float) * (op_sets.empty() ? 1 : op_sets.size()), op_sets.empty() ? nullptr : op_sets.data());
	DevBuf df(sizeof(float) * (nset ? nset : 1), f), dq(sizeof(float) * (nset ? nset : 1), Q), din(sizeof(float) * n, in), dout(sizeof(float) * n), dc(sizeof(float) * 5);
	kb_prim_filter_kernel<<<1, 32>>>(kind, nset, df.as<float>(), Q ? dq.as<float>() : nullptr, F, n, din.as<float>(), dout.as<float>(), dc.as<float>(), op, hc, op_sets.empty() ? nullptr : dsets.as<float>());
	int rc = prim_finish("kb_prim_filter"); if (rc) return rc;
	cudaMemcpy(out, dout.p, sizeof(float) * n, cudaMemcpyDeviceToHost);
	cudaMemcpy(coeffs, dc.p, sizeof(float) * 5, cudaMemcpyDeviceToHost);
	return KB_OK;
}
static int prim_env(KbEnv e, const KbFs& F, int n, int release_at, float rt, float rl, int adsr, float* out, int* stage_out) {
	DevBuf dout(sizeof(float) * n), dst(sizeof(int) * n);
	kb_prim_env_kernel<<<1, 32>>>(e, F, n, release_at, rt, rl, adsr, dout.as<float>(), dst.as<int>());
	int rc = prim_finish("kb_prim_envelope"); if (rc) return rc;
	cudaMemcpy(out, dout.p, sizeof(float) * n, cudaMemcpyDeviceToHost);
	if (stage_out) cudaMemcpy(stage_out, dst.p, sizeof(int) * n, cudaMemcpyDeviceToHost);
	return KB_OK;
}
extern "C" int kb_prim_envelope(int npts, const float* xy, int loop_start, int loop_end, float fs, int n, int release_at,
                                float release_time, float release_level, float* out, int* stage_out) {
	if (kb_device_count() < 1) return kb_fail(KB_ENODEV, "no CUDA device");
	if (npts < 0 || npts > KB_ENV_MAXPTS || !xy || !out) return kb_fail(KB_EINVAL, "kb_prim_envelope: bad argument");
	const KbFs F = kb_make_fs(fs);
	KbEnv e; kb_env_construct(F, e); kb_env_set_points(F, e, npts, xy);
	if (loop_start >= 0) kb_env_set_loop(e, loop_start, loop_end);
	return prim_env(e, F, n, release_at, release_time, release_level, 0, out, stage_out);
}
extern "C" int kb_prim_adsr(float A, float D, float S, float R, float fs, int n, int release_at, float* out, int* stage_out) {
	if (kb_device_count() < 1) return kb_fail(KB_ENODEV, "no CUDA device");
	if (!out) return kb_fail(KB_EINVAL, "kb_prim_adsr: bad argument");
	const KbFs F = kb_make_fs(fs);
	KbEnv e; kb_adsr_construct(F, e); kb_adsr_set(F, e, A, D, S, R);
	return prim_env(e, F, n, release_at, 0.f, 0.f, 1, out, stage_out);
}
extern "C" int kb_prim_math(int fn, int n, const float* x, float* out) {
	if (kb_device_count() < 1) return kb_fail(KB_ENODEV, "no CUDA device");
	if (fn < 0 || fn > 3 || !x || !out || n < 0) return kb_fail(KB_EINVAL, "kb_prim_math: bad argument");
	DevBuf dx(sizeof(float) * n, x), dout(sizeof(float) * n);
	kb_prim_math_kernel<<<148, 256>>>(fn, n, dx.as<float>(), dout.as<float>());
	int rc = prim_finish("kb_prim_math"); if (rc) return rc;
	cudaMemcpy(out, dout.p, sizeof(float) * n, cudaMemcpyDeviceToHost);
	return KB_OK;
}

// klang-b200 — sm_100a kernels of the hot path.
//
// Layouts (all float32, planar):
//   effect io       [instances][channels][n]            in place
//   voice scratch   [instances*voices][channels][n]     per-voice streams (each voice rendered alone)
//   synth out       [instances][channels][n]            Synth::process output per instance
// Voice / instance state is an array of POD blobs (kb_state.h) in HBM; a kernel loads a blob into registers /
// local memory, runs the block, and stores it back.
#pragma once
#include <cuda_runtime.h>

#include "kb_graphs.cuh"
#include "kb_fx_parallel.cuh"

// per-instance, per-block constants computed by the host at control rate (see kb_graphs.cuh)
struct KbSynthBlock { KbTbBlock tb; float sx_tr_at, sx_dt_at; float fm_i1, fm_i2; /* FM.k:63-64: controls[1], controls[2] */
                      float c[3]; /* Modulation/{AM,FM,FM2}.k: controls[0..2] */ };

// ============================================================================================ synth voices
// One lane = one voice (Note::process(buffer), klang.h:4295-4303): the lane runs the block's n-step recurrence
// with its state in registers and parks each sample in a [32 x 33] shared tile; every 32 steps the warp
// transposes the tile out so that HBM stores are 128-byte coalesced rows of one voice.
template <int GRAPH, class VOICE>
__global__ void __launch_bounds__(128) kb_voice_kernel(VOICE* __restrict__ voices, KbVoiceHdr* __restrict__ hdr,
                                                       const KbSynthBlock* __restrict__ blk, float* __restrict__ dst,
                                                       int n, int voices_per_inst, int total, KbFs fs) {
	__shared__ float tile[4][32][33];
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const int v = blockIdx.x * blockDim.x + threadIdx.x;
	const int vbase = v - lane;
	const bool valid = v < total;
	int stage = KB_NOTE_OFF;
	if (valid) stage = hdr[v].stage;
	const bool active = valid && stage != KB_NOTE_OFF;
	if (valid) hdr[v].active = active ? 1 : 0;
	if (!__any_sync(0xffffffffu, active)) {
		for (int r = 0; r < 32 && vbase + r < total; r++)
			for (int t = lane; t < n; t += 32) dst[(size_t)(vbase + r) * n + t] = 0.f;
		return;
	}
	VOICE s;
	KbTbBlock tb;
	float fm_i1 = 0.f, fm_i2 = 0.f, sm_c2 = 0.f;
	if (active) {
		s = voices[v];
		if (GRAPH == KB_SY_TB303) tb = blk[v / voices_per_inst].tb;
		if (GRAPH == KB_SY_FM) { fm_i1 = blk[v / voices_per_inst].fm_i1; fm_i2 = blk[v / voices_per_inst].fm_i2; }
		if (GRAPH == KB_SY_AM) { fm_i1 = blk[v / voices_per_inst].c[0]; fm_i2 = blk[v / voices_per_inst].c[1]; sm_c2 = blk[v / voices_per_inst].c[2]; }
	}
	for (int base = 0; base < n; base += 32) {
		const int steps = min(32, n - base);
		for (int t = 0; t < steps; t++) {
			float y = 0.f;
			if (active) {
				if constexpr (GRAPH == KB_SY_SUBTRACTIVE) y = kb_sub_tick(fs, s, stage);
				if constexpr (GRAPH == KB_SY_SUPERSAW) y = kb_ssaw_tick(fs, s, stage);
				if constexpr (GRAPH == KB_SY_TB303) y = kb_tb_tick(fs, tb, s, stage);
				if constexpr (GRAPH == KB_SY_FM) y = kb_fm_tick(fs, fm_i1, fm_i2, s, stage);
				if constexpr (GRAPH == KB_SY_BREAKPOINT) y = kb_senv_tick(fs, s, stage);         // Breakpoint.k, Ramp.k and Release.k share the voice
				if constexpr (GRAPH == KB_SY_ADDITIVE_SAW) y = kb_add_tick(fs, s);               // Additive/Saw.k and Square.k share the voice
				if constexpr (GRAPH == KB_SY_AM) y = kb_smod_tick(fs, fm_i1, fm_i2, sm_c2, s, stage);   // Modulation/AM.k, FM.k and FM2.k share the voice
			}
			tile[warp][lane][t] = y;
		}
		__syncwarp();
		for (int r = 0; r < 32 && vbase + r < total; r++)
			if (lane < steps) dst[(size_t)(vbase + r) * n + base + lane] = tile[warp][r][lane];
		__syncwarp();
	}
	if (active) {
		voices[v] = s;
		hdr[v].stage = stage;
	}
}

// Additive/Saw.k, Square.k, time-parallel: nothing in the voice is a recurrence (32 integer phase ramps), so thread = (voice, sample)
// evaluates the whole partial sum of one sample from the block-start state (kb_add_at) and a second kernel advances the phases.
// blockIdx.y = voice; the voice state is staged in shared memory.
__global__ void __launch_bounds__(256) kb_additive_kernel(const KbAddVoice* __restrict__ voices, KbVoiceHdr* __restrict__ hdr, float* __restrict__ dst, int n, KbFs fs) {
	__shared__ KbAddVoice s;
	const int v = blockIdx.y;
	const bool active = hdr[v].stage != KB_NOTE_OFF;
	if (threadIdx.x == 0 && blockIdx.x == 0) hdr[v].active = active ? 1 : 0;
	if (active) for (int w = threadIdx.x; w < (int)(sizeof(KbAddVoice) / 4); w += blockDim.x) reinterpret_cast<unsigned*>(&s)[w] = reinterpret_cast<const unsigned*>(voices + v)[w];
	__syncthreads();
	for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) dst[(size_t)v * n + t] = active ? kb_add_at(fs, s, (uint32_t)t) : 0.f;
}
__global__ void kb_additive_advance_kernel(KbAddVoice* __restrict__ voices, const KbVoiceHdr* __restrict__ hdr, int total, int n, KbFs fs) {
	const int i = blockIdx.x * blockDim.x + threadIdx.x, v = i >> 5, o = i & 31;
	if (v >= total || hdr[v].stage == KB_NOTE_OFF) return;
	if (kb_add_partial_on(fs, voices[v], o)) voices[v].osc[o].position += (uint32_t)n * (uint32_t)voices[v].osc[o].increment;
}

// Synth::process voice loop + mix (klang.h:4450-4456 / 4842-4848): thread = (instance, sample).  Voices are
// combined sequentially in voice-index order, exactly as the reference sweeps them: mono Note::process
// ASSIGNS (the last active voice wins), stereo / KB_MIX_SUM accumulates in fp32.
__global__ void __launch_bounds__(256) kb_mix_kernel(const float* __restrict__ scratch, const KbVoiceHdr* __restrict__ hdr, float* __restrict__ out,
                                                     int n, int voices, int sum_mode) {
	__shared__ int s_active[KB_MAX_VOICES];
	const int inst = blockIdx.y;
	for (int v = threadIdx.x; v < voices; v += blockDim.x) s_active[v] = hdr[(size_t)inst * voices + v].active;
	__syncthreads();
	const int t = blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= n) return;
	float acc = 0.f;
	const float* s = scratch + (size_t)inst * voices * n + t;
	if (!sum_mode) {
		// mono Note::process assigns (klang.h:4299): the block is whatever the last active voice wrote
		int last = -1;
		for (int v = voices - 1; v >= 0 && last < 0; v--) if (s_active[v]) last = v;
		out[(size_t)inst * n + t] = last >= 0 ? __ldcs(s + (size_t)last * n) : 0.f;
		return;
	}
	for (int v0 = 0; v0 < voices; v0 += 8) {
		float x[8];
		#pragma unroll
		for (int j = 0; j < 8; j++) x[j] = (v0 + j < voices && s_active[v0 + j]) ? __ldcs(s + (size_t)(v0 + j) * n) : 0.f;   // 8 loads in flight
		#pragma unroll
		for (int j = 0; j < 8; j++) if (v0 + j < voices && s_active[v0 + j]) acc = acc + x[j];                                // summed in voice order
	}
	out[(size_t)inst * n + t] = acc;
}

// The same voice sum for a whole bank in one launch (round 2; sum mode, one channel row per voice): CTA = 32 samples of every instance.  All
// threads first pull the CTA's [instances][voices][32] slab of the per-voice streams from L2 with every load in flight at once (the loop of
// kb_mix_kernel walks the 128 voices eight loads at a time: 16 dependent L2 round trips); then one warp per instance adds its voices in voice
// order from shared memory (the reference's order, klang.h:4450-4456 / 4842-4848) and writes the instance output; warp 0 finally adds the
// instance sums in instance order into the bank mix (kb_bank_mix_kernel's sum).  Instances are taken in groups that fit the shared memory.
#define KB_MIXF_TS 32
__global__ void __launch_bounds__(1024) kb_mix_fused_kernel(const float* __restrict__ scratch, const KbVoiceHdr* __restrict__ hdr, float* __restrict__ inst_out,
                                                            float* __restrict__ bank_out, int n, int voices, int instances, int group) {
	extern __shared__ __align__(16) unsigned char kb_mixf_smem[];
	float* tile = reinterpret_cast<float*>(kb_mixf_smem);                   // [group][voices][32]
	float* isum = tile + (size_t)group * voices * KB_MIXF_TS;               // [group][32]
	int* s_act = reinterpret_cast<int*>(isum + (size_t)group * KB_MIXF_TS);  // [group][voices]
	const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nwarps = blockDim.x >> 5;
	const int t = blockIdx.x * KB_MIXF_TS + lane;
	float bank = 0.f;
	// launched as the programmatic dependent of the voice kernel (kb_api.cu): this CTA may already be resident while the last voice CTAs
	// run; everything the voice kernel wrote is visible once the wait returns (a plain launch passes straight through)
	asm volatile("griddepcontrol.wait;" ::: "memory");
	for (int i0 = 0; i0 < instances; i0 += group) {
		const int gi = min(group, instances - i0), rows = gi * voices;
		if (i0) __syncthreads();
		for (int r = tid; r < rows; r += blockDim.x) s_act[r] = hdr[(size_t)i0 * voices + r].active;
		__syncthreads();
		// full 32-sample groups with 16-byte aligned rows: a warp pulls FOUR rows per instruction (8 lanes x 16 bytes per row) — a quarter of the
		// load / store instructions of the lane = sample form below, which remains for ragged blocks
		const int t0 = blockIdx.x * KB_MIXF_TS;
		const bool vec = (n & 3) == 0 && t0 + KB_MIXF_TS <= n && (reinterpret_cast<size_t>(scratch) & 15) == 0;
		if (vec) {
			const int sub = lane >> 3, c4 = (lane & 7) * 4;
			for (int g0 = warp; g0 * 4 < rows; g0 += 4 * nwarps) {
				float4 x[4];
				#pragma unroll
				for (int j = 0; j < 4; j++) {
					const int r = (g0 + j * nwarps) * 4 + sub;
					x[j] = (r < rows && s_act[r < rows ? r : 0]) ? __ldcs(reinterpret_cast<const float4*>(scratch + ((size_t)i0 * voices + r) * n + t0 + c4)) : make_float4(0.f, 0.f, 0.f, 0.f);
				}
				#pragma unroll
				for (int j = 0; j < 4; j++) { const int r = (g0 + j * nwarps) * 4 + sub; if (r < rows) *reinterpret_cast<float4*>(&tile[(size_t)r * KB_MIXF_TS + c4]) = x[j]; }
			}
		} else
		for (int r0 = warp; r0 < rows; r0 += 16 * nwarps) {
			float x[16];
			#pragma unroll
			for (int j = 0; j < 16; j++) {
				const int r = r0 + j * nwarps;
				x[j] = (r < rows && t < n && s_act[r < rows ? r : 0]) ? __ldcs(scratch + ((size_t)i0 * voices + r) * n + t) : 0.f;
			}
			#pragma unroll
			for (int j = 0; j < 16; j++) { const int r = r0 + j * nwarps; if (r < rows) tile[(size_t)r * KB_MIXF_TS + lane] = x[j]; }
		}
		__syncthreads();
		for (int i = warp; i < gi; i += nwarps) {
			const float* tp = tile + (size_t)i * voices * KB_MIXF_TS + lane;
			const int* ap = s_act + i * voices;
			float acc = 0.f;
			#pragma unroll 16                                                   // (operands of 16 voices in flight; the adds stay in voice order)
			for (int v = 0; v < voices; v++) if (ap[v]) acc = acc + tp[(size_t)v * KB_MIXF_TS];   // summed in voice order; inactive voices are skipped like kb_mix_kernel
			isum[i * KB_MIXF_TS + lane] = acc;
			if (t < n) inst_out[(size_t)(i0 + i) * n + t] = acc;
		}
		__syncthreads();
		if (warp == 0 && bank_out) for (int i = 0; i < gi; i++) bank += isum[i * KB_MIXF_TS + lane];
	}
	if (warp == 0 && bank_out && t < n) bank_out[t] = bank;
}

// event upload: dirty voices travel packed in one staging buffer [count][hdr | blob] and are scattered to their slots
__global__ void kb_scatter_voices_kernel(const unsigned char* __restrict__ staging, const int* __restrict__ index, int count, int voice_bytes,
                                         KbVoiceHdr* __restrict__ hdr, unsigned char* __restrict__ vstate) {
	const int rec_words = (int)(sizeof(KbVoiceHdr) + voice_bytes) / 4, hdr_words = (int)sizeof(KbVoiceHdr) / 4;
	for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < count * rec_words; i += gridDim.x * blockDim.x) {
		const int k = i / rec_words, w = i % rec_words, v = index[k];
		const unsigned val = reinterpret_cast<const unsigned*>(staging)[i];
		if (w < hdr_words) reinterpret_cast<unsigned*>(hdr + v)[w] = val;
		else reinterpret_cast<unsigned*>(vstate + (size_t)v * voice_bytes)[w - hdr_words] = val;
	}
}

// Sum of the per-instance outputs in instance order: the bank mix [channels][n] that is reduced across GPUs.
__global__ void kb_bank_mix_kernel(const float* __restrict__ inst_out, float* __restrict__ out, int cn, int instances) {
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= cn) return;
	float acc = 0.f;
	for (int k = 0; k < instances; k++) acc += inst_out[(size_t)k * cn + i];
	out[i] = acc;
}

// ------------------------------------------------------------------------------------------------- SynTHX
// SynTHX voices are 88-132 band-limited saws each; all of them are pure functions of an integer phase ramp
// (kb_osm_at), so the graph is evaluated with thread = output sample: each thread walks voices, additive notes,
// harmonics and partials in exactly the reference's accumulation order (SynTHX.k:109-122, 170-182) and keeps
// the running buffer value in a register.  The only true recurrence — the ADSR — is swept first by a
// lane-per-voice kernel into adsr[voice][n].

// once per block, thread = (voice, partial): Partial::set(transpose, detune) (SynTHX.k:37-45, 75-81)
__global__ void kb_sx_prepare_kernel(KbSxVoice* __restrict__ voices, const KbVoiceHdr* __restrict__ hdr, const KbSynthBlock* __restrict__ blk,
                                     int voices_per_inst, int total, KbFs fs) {
	const int idx = blockIdx.x * blockDim.x + threadIdx.x;
	const int v = idx / 132, p = idx % 132;
	if (v >= total || hdr[v].stage == KB_NOTE_OFF) return;
	const KbSynthBlock& b = blk[v / voices_per_inst];
	KbSxPartial& P = voices[v].notes[p / 12].partial[(p % 12) / 3][p % 3];
	kb_sx_partial_set2(fs, P, b.sx_tr_at, b.sx_dt_at);
}
// lane = voice: ADSR sweep (SynTHX.k:176-178) and stop() (:179-180)
__global__ void kb_sx_adsr_kernel(KbSxVoice* __restrict__ voices, KbVoiceHdr* __restrict__ hdr, float* __restrict__ adsr, int n, int total, KbFs fs) {
	const int v = blockIdx.x * blockDim.x + threadIdx.x;
	if (v >= total) return;
	const int stage = hdr[v].stage;
	hdr[v].active = stage != KB_NOTE_OFF;
	if (stage == KB_NOTE_OFF) return;
	KbEnvR e;
	kb_envr_load(e, voices[v].adsr);
	kb_envr_run(fs, e, voices[v].adsr.px, voices[v].adsr.py, adsr + (size_t)v * n, n);
	kb_envr_store(e, voices[v].adsr);
	if (e.stage == KB_ENV_OFF) hdr[v].stage = KB_NOTE_OFF;
}
// Render.  One CTA = (instance, 32 consecutive samples): the running buffer values of those samples live in the registers
// of ONE consumer warp (lane = sample), which performs the reference's additions strictly in order; 24 producer warps —
// two per partial [harmonic][partial] pair, splitting the 11 Additive notes — evaluate the oscillators of the NEXT voice
// (closed form, lane = sample) into a double-buffered shared tile while the consumer adds the current one.  One round =
// one voice = 132 partial slots, one __syncthreads.  The fp32 sum therefore has exactly the reference's association, while
// the oscillator evaluation runs in parallel on all warps.
// per_voice: every voice starts from a cleared buffer and is written to dst[voice][2][n]; otherwise the instance buffer
// is carried from voice to voice (each voice's `buffer *= adsr` also scales what earlier voices left there, exactly as
// Stereo::Synth::process hands the same buffer to every note, klang.h:4842-4848) and the synth-level tanh post-fx
// (SynTHX.k:201-206) is applied at the end.
#define KB_SX_PRODUCERS 12                  // (harmonic, partial) pairs of an Additive note
#define KB_SX_WARPS_PER_PAIR 2              // producer warps sharing one pair: they split the 11 notes
struct KbSxSmem {
	float val[2][11][KB_SX_PRODUCERS][32];
	int flag[2][11][KB_SX_PRODUCERS];          // bit 0: slot is ticked this block, bit 1: routed to the right channel
	unsigned voice[2][sizeof(KbSxAdditive) * 11 / 4];   // oscillator state of the current and the next voice
	int vlist[KB_MAX_VOICES], nact;
};
__global__ void __launch_bounds__(32 * (KB_SX_PRODUCERS * KB_SX_WARPS_PER_PAIR + 1)) kb_sx_render_kernel(const KbSxVoice* __restrict__ voices, const KbVoiceHdr* __restrict__ hdr,
                                                                                const float* __restrict__ adsr, float* __restrict__ dst, int n,
                                                                                int voices_per_inst, int per_voice) {
	constexpr int NT = 32 * (KB_SX_PRODUCERS * KB_SX_WARPS_PER_PAIR + 1);
	constexpr int VW = (int)(sizeof(KbSxAdditive) * 11 / 4);       // words of a voice's 11 Additive notes (the ADSR follows them)
	constexpr int VPT = (VW + NT - 1) / NT;                          // words per thread when a voice is staged
	extern __shared__ __align__(16) unsigned char kb_sx_smem_raw[];
	KbSxSmem& S = *reinterpret_cast<KbSxSmem*>(kb_sx_smem_raw);
	const int inst = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const int t = blockIdx.x * 32 + lane;
	if (threadIdx.x == 0) {
		int c = 0;
		for (int vi = 0; vi < voices_per_inst; vi++) if (hdr[inst * voices_per_inst + vi].active) S.vlist[c++] = inst * voices_per_inst + vi;
		S.nact = c;
	}
	__syncthreads();
	const int nact = S.nact;
	float l = 0.f, r = 0.f;
	if (per_voice && warp == 0) {      // voices that are Off this block: cleared streams
		for (int vi = 0; vi < voices_per_inst; vi++) {
			const int v = inst * voices_per_inst + vi;
			if (!hdr[v].active && t < n) { dst[((size_t)v * 2 + 0) * n + t] = 0.f; dst[((size_t)v * 2 + 1) * n + t] = 0.f; }
		}
	}
	if (nact > 0) {
		const unsigned* src = reinterpret_cast<const unsigned*>(&voices[S.vlist[0]]);
		for (int w = threadIdx.x; w < VW; w += NT) S.voice[0][w] = src[w];
	}
	__syncthreads();
	unsigned pre[VPT];                                 // the next voice travels through registers during the current one
	for (int j = 0; j <= nact; j++) {
		if (j + 1 < nact) {
			const unsigned* src = reinterpret_cast<const unsigned*>(&voices[S.vlist[j + 1]]);
			#pragma unroll
			for (int q = 0; q < VPT; q++) { const int w = threadIdx.x + q * NT; if (w < VW) pre[q] = src[w]; }
		}
		if (warp > 0 && j < nact) {
			// ---- producers: voice j; this warp's partial of each of the 11 notes
			const KbSxAdditive* notes = reinterpret_cast<const KbSxAdditive*>(S.voice[j & 1]);
			const int w = (warp - 1) % KB_SX_PRODUCERS, p = w / 3, q = w % 3;
			for (int k = (warp - 1) / KB_SX_PRODUCERS; k < 11; k += KB_SX_WARPS_PER_PAIR) {
				const KbSxAdditive& A = notes[k];
				int flag = 0;
				float x = 0.f;
				if (q < ((A.frequency < 440.f) ? 2 : 3)) {                // Additive::process picks partials<2> or <3>  SynTHX.k:124-127
					const KbSxPartial& P = A.partial[p][q];
					x = kb_osm_at(P.osc, (uint32_t)t) * 0.25f;            // partial * 0.25f  SynTHX.k:113-118
					flag = 1 | (P.right ? 2 : 0);
				}
				S.val[j & 1][k][w][lane] = x;
				if (lane == 0) S.flag[j & 1][k][w] = flag;
			}
		} else if (warp == 0 && j > 0) {
			// ---- consumer: voice j-1, additions in the reference's order (note, harmonic, partial)
			const int b = (j - 1) & 1;
			const int v = S.vlist[j - 1];
			const float a = t < n ? adsr[(size_t)v * n + t] : 0.f;          // (issued before the additions: off the critical path)
			#pragma unroll
			for (int k = 0; k < 11; k++) {
				#pragma unroll
				for (int w = 0; w < KB_SX_PRODUCERS; w++) {
					// branch-free: the addition happens exactly when the reference performs it (a skipped slot must not add 0.0)
					const int f = S.flag[b][k][w];
					const float x = S.val[b][k][w][lane];
					const float nl = l + x, nr = r + x;
					l = (f == 1) ? nl : l;
					r = (f == 3) ? nr : r;
				}
			}
			l *= a; r *= a;                                               // the voice is complete: buffer *= adsr++  SynTHX.k:176-178
			if (per_voice) {
				if (t < n) { dst[((size_t)v * 2 + 0) * n + t] = l; dst[((size_t)v * 2 + 1) * n + t] = r; }
				l = 0.f; r = 0.f;
			}
		}
		if (j + 1 < nact) {
			#pragma unroll
			for (int q = 0; q < VPT; q++) { const int w = threadIdx.x + q * NT; if (w < VW) S.voice[(j + 1) & 1][w] = pre[q]; }
		}
		__syncthreads();
	}
	if (!per_voice && warp == 0 && t < n) {
		const float gain = 0.5f, _tanh = 0.761594155956f;
		dst[((size_t)inst * 2 + 0) * n + t] = kb_tanhf(l * gain) * _tanh;
		dst[((size_t)inst * 2 + 1) * n + t] = kb_tanhf(r * gain) * _tanh;
	}
}
// thread = (voice, partial): advance every oscillator by the block's n ticks
__global__ void kb_sx_advance_kernel(KbSxVoice* __restrict__ voices, const KbVoiceHdr* __restrict__ hdr, int n, int total) {
	const int idx = blockIdx.x * blockDim.x + threadIdx.x;
	const int v = idx / 132, p = idx % 132;
	if (v >= total || !hdr[v].active) return;
	KbSxAdditive& A = voices[v].notes[p / 12];
	const int q = p % 3;
	if (q >= ((A.frequency < 440.f) ? 2 : 3)) return;
	kb_osm_advance(A.partial[(p % 12) / 3][q].osc, (uint32_t)n);
}

// ================================================================================================ effects
// Streaming rows in place, 16-byte vectors, FOUR loads in flight per thread before the first store (round 2: with one vector in flight per
// thread and a grid of one wave plus a few CTAs — the few ran alone, latency-bound, after the wave — Gain.k reached 0.74 of the copy peak).
// The launch covers a row with one trip per thread (kb_stream_grid_x in kb_api.cu): many short CTAs, no tail.  `f(x, t)` maps sample t of the row.
template <int U = 4, class F>
KB_D void kb_stream_row(float* __restrict__ p, int n, F f) {
	const int n4 = ((reinterpret_cast<uintptr_t>(p) & 15) == 0) ? (n >> 2) : 0;
	float4* p4 = reinterpret_cast<float4*>(p);
	const int stride = gridDim.x * blockDim.x;
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	auto map4 = [&](float4 x, int i4) { const uint32_t t = (uint32_t)i4 << 2; x.x = f(x.x, t); x.y = f(x.y, t + 1); x.z = f(x.z, t + 2); x.w = f(x.w, t + 3); return x; };
	if (U == 4) {
		for (; i + 3 * stride < n4; i += 4 * stride) {
			float4 x0 = __ldcs(p4 + i), x1 = __ldcs(p4 + i + stride), x2 = __ldcs(p4 + i + 2 * stride), x3 = __ldcs(p4 + i + 3 * stride);
			__stcs(p4 + i, map4(x0, i)); __stcs(p4 + i + stride, map4(x1, i + stride));
			__stcs(p4 + i + 2 * stride, map4(x2, i + 2 * stride)); __stcs(p4 + i + 3 * stride, map4(x3, i + 3 * stride));
		}
	}
	#pragma unroll 1
	for (; i < n4; i += stride) __stcs(p4 + i, map4(__ldcs(p4 + i), i));
	#pragma unroll 1
	for (int j = (n4 << 2) + blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride) p[j] = f(p[j], (uint32_t)j);
}
// Sweeps whose body gathers from a delay ring: ONE frame per thread and trip, so that the lanes of a warp tap consecutive ring slots (with
// four consecutive frames per thread every tap instruction of a warp touches four times as many lines: measured slower for Flanger.k),
// U independent trips in flight.  STORE = false: `f(x, t)` only consumes sample t.
template <int U, bool STORE, class F>
KB_D void kb_sweep_row(float* __restrict__ p, int n, F f) {
	const int stride = gridDim.x * blockDim.x;
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	for (; i + (U - 1) * stride < n; i += U * stride) {
		float x[U];
		#pragma unroll
		for (int u = 0; u < U; u++) x[u] = p[i + u * stride];
		#pragma unroll
		for (int u = 0; u < U; u++) { const float y = f(x[u], (uint32_t)(i + u * stride)); if (STORE) p[i + u * stride] = y; }
	}
	#pragma unroll 1
	for (; i < n; i += stride) { const float y = f(p[i], (uint32_t)i); if (STORE) p[i] = y; }
}
// Gain.k: no state, one multiply per sample.
__global__ void __launch_bounds__(256) kb_gain_kernel(const KbFxHdr* __restrict__ hdr, float* __restrict__ io, int n) {
	const int inst = blockIdx.y;
	const float gain = hdr[inst].controls[0].value;
	kb_stream_row(io + (size_t)inst * n, n, [gain](float x, uint32_t) { return x * gain; });
}

// Pan.k / RM.k / Tremolo.k / Clipping.k: elementwise streaming kernel over planar rows [instance][channel][n]; blockIdx.y = instance * channels +
// channel.  The LFO phase of sample t is closed-form (kb_ew_sample).
template <int GRAPH>
__global__ void __launch_bounds__(256) kb_elementwise_kernel(int channels, const KbFxHdr* __restrict__ hdr, const KbLfoFx* __restrict__ lfos,
                                      float* __restrict__ io, int n, int stride) {
	constexpr int graph = GRAPH;                                           // (a compile-time graph: no per-sample dispatch in kb_ew_sample)
	const int inst = blockIdx.y / channels, ch = blockIdx.y % channels;
	const float c0 = hdr[inst].controls[0].value, c1 = hdr[inst].controls[1].value;
	KbFastSine lfo = { 0.f, 0, 0u, 0u };
	if (lfos) lfo = lfos[inst].lfo;
	kb_stream_row(io + ((size_t)inst * channels + ch) * stride, n, [=](float x, uint32_t t) { return kb_ew_sample(graph, c0, c1, lfo, ch, t, x); });
}
// after the block: the n ticks the LFO made (Fast::Sine::process, klang.h:5164-5170)
__global__ void kb_lfo_advance_kernel(KbLfoFx* __restrict__ lfos, int instances, int n) {
	const int inst = blockIdx.x * blockDim.x + threadIdx.x;
	if (inst < instances) lfos[inst].lfo.position += (uint32_t)n * (uint32_t)lfos[inst].lfo.increment;
}

// Echo.k, time-parallel (kb_echo_write_at / kb_echo_read_at): two streaming sweeps over the block and the position advance; blockIdx.y = instance
__global__ void kb_echo_write_kernel(const KbOneDelayFx* __restrict__ st, float* __restrict__ rings, const float* __restrict__ io, int n, int stride) {
	const KbOneDelayFx s = st[blockIdx.y];
	const float* p = io + (size_t)blockIdx.y * stride;
	kb_sweep_row<4, false>(const_cast<float*>(p), n, [&](float x, uint32_t t) { kb_echo_write_at(s, rings, (int)t, x); return 0.f; });
}
__global__ void kb_echo_read_kernel(const KbFxHdr* __restrict__ hdr, const KbOneDelayFx* __restrict__ st, const float* __restrict__ rings,
                                    float* __restrict__ io, int n, int stride, KbFs fs) {
	const KbOneDelayFx s = st[blockIdx.y];
	const KbFxHdr& h = hdr[blockIdx.y];
	float* p = io + (size_t)blockIdx.y * stride;
	kb_sweep_row<4, true>(p, n, [&](float x, uint32_t t) { return kb_echo_read_at(fs, h, s, rings, (int)t, x); });
}
// Flanger.k / Modulation/Chorus.k, time-parallel (kb_modline_*): LFO settings of the block's first frame, write sweep with stash, read sweep,
// LFO / position advance.  blockIdx.y = instance; `old` is [instances][stride] scratch.
__global__ void kb_modline_begin_kernel(int graph, KbFxHdr* __restrict__ hdr, KbModDelayFx* __restrict__ st, int instances, KbFs fs, float* __restrict__ depth, int n, int stride) {
	const int inst = blockIdx.x * blockDim.x + threadIdx.x;
	if (inst < instances) kb_modline_begin(graph, fs, hdr[inst], st[inst], depth ? depth + (size_t)inst * stride : nullptr, n);
}
__global__ void kb_modline_write_kernel(const KbModDelayFx* __restrict__ st, float* __restrict__ rings, float* __restrict__ old, const float* __restrict__ io, int n, int stride) {
	const KbDelay d = st[blockIdx.y].delay;
	const float* p = io + (size_t)blockIdx.y * stride;
	float* o = old + (size_t)blockIdx.y * stride;
	kb_sweep_row<4, false>(const_cast<float*>(p), n, [&](float x, uint32_t t) { kb_modline_write_at(d, rings, o, (int)t, x); return 0.f; });
}
__global__ void kb_modline_read_kernel(int graph, const KbFxHdr* __restrict__ hdr, const KbModDelayFx* __restrict__ st, const float* __restrict__ rings,
                                       const float* __restrict__ old, const float* __restrict__ depth, float* __restrict__ io, int n, int stride, KbFs fs) {
	__shared__ KbModDelayFx s;
	__shared__ KbFxHdr h;
	for (int w = threadIdx.x; w < (int)(sizeof(KbModDelayFx) / 4); w += blockDim.x) reinterpret_cast<unsigned*>(&s)[w] = reinterpret_cast<const unsigned*>(st + blockIdx.y)[w];
	for (int w = threadIdx.x; w < (int)(sizeof(KbFxHdr) / 4); w += blockDim.x) reinterpret_cast<unsigned*>(&h)[w] = reinterpret_cast<const unsigned*>(hdr + blockIdx.y)[w];
	__syncthreads();
	float* p = io + (size_t)blockIdx.y * stride;
	const float* o = old + (size_t)blockIdx.y * stride;
	const float* dr = depth ? depth + (size_t)blockIdx.y * stride : nullptr;
	kb_sweep_row<2, true>(p, n, [&](float x, uint32_t t) { return kb_modline_read_at(graph, fs, h, s, rings, o, dr, n, (int)t, x); });
}
__global__ void kb_modline_end_kernel(int graph, KbModDelayFx* __restrict__ st, int instances, int n) {
	const int inst = blockIdx.x * blockDim.x + threadIdx.x;
	if (inst < instances) kb_modline_end(graph, st[inst], n);
}
// Feedback.k, chunk-parallel (kb_feedback_chunk / kb_feedback_at): one CTA per instance walks the block chunk by chunk, one barrier per chunk;
// instances whose delay is too short for a chunk (chunk 0) run frame by frame on thread 0.  The position advances at the end.
__global__ void __launch_bounds__(1024) kb_feedback_par_kernel(const KbFxHdr* __restrict__ hdr, KbOneDelayFx* __restrict__ st, float* __restrict__ rings,
                                                                float* __restrict__ io, int n, int stride, KbFs fs) {
	const KbFxHdr& h = hdr[blockIdx.x];
	const KbOneDelayFx s = st[blockIdx.x];
	float* p = io + (size_t)blockIdx.x * stride;
	const int chunk = kb_feedback_chunk(fs, n, h.controls[0].value);
	if (chunk == 0) {
		if (threadIdx.x == 0) for (int t = 0; t < n; t++) p[t] = kb_feedback_at(fs, h, s, rings, t, p[t]);
	} else {
		for (int c0 = 0; c0 < n; c0 += chunk) {
			const int c1 = min(n, c0 + chunk);
			for (int t = c0 + threadIdx.x; t < c1; t += blockDim.x) p[t] = kb_feedback_at(fs, h, s, rings, t, p[t]);
			__syncthreads();                                                 // the chunk's ring writes are visible to the next chunk's taps
		}
	}
	__syncthreads();
	if (threadIdx.x == 0) st[blockIdx.x].delay.position = (s.delay.position + n) % s.delay.SIZE;
}
__global__ void kb_onedelay_advance_kernel(KbOneDelayFx* __restrict__ st, int instances, int n) {
	const int inst = blockIdx.x * blockDim.x + threadIdx.x;
	if (inst < instances) st[inst].delay.position = (st[inst].delay.position + n) % st[inst].delay.SIZE;
}

// Delay-line effects, sequential form: one lane = one instance (Effect::process(buffer), klang.h:4208-4216 /
// 4708-4716), rings in HBM.  Exact for any control setting; the time-parallel kernels below take over whenever
// the feedback delays allow.
template <int GRAPH, class STATE>
__global__ void kb_fx_seq_kernel(KbFxHdr* __restrict__ hdrs, STATE* __restrict__ states, float* __restrict__ rings,
                                 float* __restrict__ io, int n, int stride, int channels, int instances, KbFs fs, const KbFxPlan* __restrict__ plan) {
	const int inst = blockIdx.x * blockDim.x + threadIdx.x;
	if (inst >= instances) return;
	if (plan && (plan[inst].mode & KB_PLAN_PARALLEL)) return;         // taken by the chunk-parallel kernel
	KbFxHdr h = hdrs[inst];
	STATE s = states[inst];
	float* l = io + (size_t)inst * channels * stride;
	float* r = l + stride;
	for (int i = 0; i < n; i++) {
		if constexpr (GRAPH == KB_FX_PINGPONG) { float ol, orr; kb_pingpong_frame(fs, h, s, rings, l[i], r[i], ol, orr); l[i] = ol; r[i] = orr; }
		if constexpr (GRAPH == KB_FX_REVERB) { float ol, orr; kb_reverb_frame(h, s, rings, l[i], r[i], ol, orr); l[i] = ol; r[i] = orr; }
		if constexpr (GRAPH == KB_FX_DELAY_PINGPONG) { float ol, orr; kb_dpingpong_frame(fs, h, s, rings, l[i], r[i], ol, orr); l[i] = ol; r[i] = orr; }
		if constexpr (GRAPH == KB_FX_DELAY_REVERB) { l[i] = kb_dreverb_frame(fs, h, s, rings, l[i]); }
		if constexpr (GRAPH == KB_FX_ECHO) { l[i] = kb_echo_frame(fs, h, s, rings, l[i]); }
		if constexpr (GRAPH == KB_FX_IIR) { l[i] = kb_iir_frame(h, s, l[i]); }
		if constexpr (GRAPH == KB_FX_WAHWAH) { l[i] = kb_wahwah_frame(fs, h, s, l[i]); }
		if constexpr (GRAPH == KB_FX_FLANGER || GRAPH == KB_FX_MODDELAY || GRAPH == KB_FX_MOD_CHORUS) { l[i] = kb_moddelay_frame(GRAPH, fs, h, s, rings, l[i]); }
		if constexpr (GRAPH == KB_FX_FEEDBACK) { l[i] = kb_feedback_frame(fs, h, s, rings, l[i]); }
	}
	hdrs[inst] = h;
	states[inst] = s;
}

// ================================================================================================ debug taps
// `x >> debug` (klang.h:3132-3287): Debug::input ADDS x into the sample of the current frame of Debug::buffer, which the host's Debug::Session
// cleared before the block and the block drivers step once per frame (klang.h:4214, 4714) — so a block's capture is `0.f + x` per frame.
// The four bound programs that tap a signal (PingPong.k:61, Gain/RM.k:22, Gain/Tremolo.k:27, Modulation/ModDelay.k:24) all tap a quantity of
// the CONTROL half of the frame (smoothers and LFOs: no audio, no ring), so the capture is produced by a replay of that half on a copy of the
// state, ahead of the block's kernels whatever schedule they run on.  Launched only while kb_fx_bank_debug_enable is on; dbg = [instances][stride].
__global__ void kb_debug_tap_kernel(int graph, const KbFxHdr* __restrict__ hdrs, const unsigned char* __restrict__ states, size_t state_bytes,
                                    float* __restrict__ dbg, int n, int stride, int instances, KbFs fs) {
	const int inst = blockIdx.x * blockDim.x + threadIdx.x;
	if (inst >= instances) return;
	KbFxHdr h = hdrs[inst];
	float* row = dbg + (size_t)inst * stride;
	if (graph == KB_FX_PINGPONG) {                                         // controls[1].smoothed >> debug   PingPong.k:61
		KbPingPong p = *reinterpret_cast<const KbPingPong*>(states + inst * state_bytes);
		for (int i = 0; i < n; i++) { float gain, delay, dry; kb_pingpong_control(fs, h, p, gain, delay, dry); row[i] = 0.f + h.controls[1].smoothed; }
	} else if (graph == KB_FX_MODDELAY) {                                  // mod * 10.f >> debug             ModDelay.k:19-24
		KbFastSine lfo = reinterpret_cast<const KbModDelayFx*>(states + inst * state_bytes)->lfo[0];
		for (int i = 0; i < n; i++) {
			const float sm = kb_control_smooth(h.controls[1]);
			const float depth = (sm * sm * sm) / 10.f;
			kb_fsine_set_f(fs, lfo, h.controls[0].value);
			const float mod = kb_fsine_tick(lfo) * depth + depth;
			row[i] = 0.f + mod * 10.f;
		}
	}
}
// RM.k / Tremolo.k: mod >> debug, the LFO in closed form like kb_ew_sample; blockIdx.y = instance
__global__ void kb_debug_tap_lfo_kernel(int graph, const KbFxHdr* __restrict__ hdr, const KbLfoFx* __restrict__ lfos, float* __restrict__ dbg, int n, int stride) {
	const KbFastSine lfo = lfos[blockIdx.y].lfo;
	const float c1 = hdr[blockIdx.y].controls[1].value;
	float* row = dbg + (size_t)blockIdx.y * stride;
	for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) {
		const float s = kb_fsine_value(lfo.position + (uint32_t)t * (uint32_t)lfo.increment + lfo.offset);
		row[t] = 0.f + (graph == KB_FX_RM ? s : s * c1 + (1 - c1));
	}
}

// ============================================================================================= primitives
__global__ void kb_prim_delay_kernel(int n, const float* in, const int* di, const float* df, const float* set_at, float* ring,
                                     float* out_i, float* out_f, float* out_p, float* out_l) {
	if (threadIdx.x || blockIdx.x) return;
	kb_delay_kat(n, in, di, df, set_at, ring, out_i, out_f, out_p, out_l);
}
// Generators::Basic::Noise / Fast::Noise (klang.h:4947-4951, 5357-5366): n ticks = n draws of the libc stream whose captured state
// arrives by value (kb_rand.h); the host advances its copy by the same n draws and hands it back to libc
__global__ void kb_prim_noise_kernel(int fast, KbRand g, int n, float* out) {
	if (threadIdx.x || blockIdx.x) return;
	for (int s = 0; s < n; s++) { const uint32_t r = kb_rand_next(g); out[s] = fast ? kb_noise_fast(r) : kb_noise_basic(r); }
}
__global__ void kb_prim_osc_kernel(int kind, int nargs, float f, float phase, float duty, KbFs fs, int n, float* out, const float* table) {
	if (threadIdx.x || blockIdx.x) return;
	if (kind <= 3) {
		const int wf[4] = { 0, 0, 1, 1 };
		const float dt[4] = { 0.f, 1.f, 1.f, 0.5f };
		KbOsm o; kb_osm_construct(o, wf[kind], dt[kind]);
		if (nargs == 1) kb_osm_set_f(fs, o, f); else if (nargs == 2) kb_osm_set_fp(fs, o, f, phase); else kb_osm_set_fpd(fs, o, f, phase, duty);
		// even samples from the sequential form, odd samples from the closed form: both must agree with the reference
		KbOsm o0 = o;
		for (int s = 0; s < n; s++) { const float y = kb_osm_tick(o); out[s] = (s & 1) ? kb_osm_at(o0, (uint32_t)s) : y; }
	} else if (kind == 4) {   // Fast::Sine  klang.h:5135-5172, fastsinp 5117-5132, polysin 5093-5096
		KbFastSine o = { 1000.f, 0, 0u, 0u };
		if (nargs >= 2) { o.position = kb_phase_from_radians(phase); o.offset = kb_phase_from_radians(0.f); }
		if (f != o.frequency) { o.frequency = f; o.increment = kb_increment_set(fs, f); }
		for (int s = 0; s < n; s++) {
			float x = (kb_bits(((o.position + o.offset) >> 9) | 0x3f800000) - 1.f) * KB_TWO_PI_F;
			if (x > 3.f / 2.f * KB_PI_F) x -= KB_TWO_PI_F; else if (x > KB_PI_F / 2.f) x = KB_PI_F - x;
			const float x2 = x * x;
			out[s] = (((-0.00018542f * x2 + 0.0083143f) * x2 - 0.16666f) * x2 + 1.0f) * x;
			o.position += (uint32_t)o.increment;
		}
	} else if (kind == 5) {   // Basic::Sine
		KbBasicOsc o; kb_bosc_init(o);
		if (nargs == 1) kb_bosc_set_f(fs, o, f); else kb_bosc_set_fp(fs, o, f, phase);
		for (int s = 0; s < n; s++) out[s] = kb_bosc_sine_tick(o);
	} else if (kind >= 6 && kind <= 9) {   // Basic::Saw / Triangle / Square / Pulse (Pulse::set(f, phase, duty), klang.h:4935-4938)
		KbBasicOsc o; kb_bosc_init(o);
		if (nargs == 1) kb_bosc_set_f(fs, o, f); else if (nargs == 2) kb_bosc_set_fp(fs, o, f, phase);
		else if (kind == 9) { kb_bosc_set_fp(fs, o, f, phase); o.duty = duty; }
		for (int s = 0; s < n; s++) out[s] = kb_bosc_shape_tick(o, kind - 5);
	} else if (kind == 10 || kind == 11) {   // Wavetable::process + buffer::operator[](float)  klang.h:3672-3675, 2070-2078
		float increment = f * (2048 / fs.f), position = (nargs >= 2) ? phase * 2048.f : 0.f;
		for (int s = 0; s < n; s++) {
			if (!(increment >= 2048.f)) { position += increment; if (position > 2048.f) position -= 2048.f; }
			const float off = position, fl = floorf(off), frac = off - fl;
			const int i = (int)off, j = (i == 2047) ? 0 : (i + 1);
			out[s] = table[i] * (1.f - frac) + table[j] * frac;
		}
	}
}
// klang::Sample (klang.h:3679-3720) over a table in device memory (size + 1 floats: the reference reads samples[size] when the position reaches
// `size` exactly — Phase::operator+= wraps on `>`, klang.h:1527-1534 — the extra slot holds 0): set(f) -> increment 1; set(f, phase) ->
// position = phase * 44100 first; per tick the position advances and the table is read with buffer::operator[](float) (klang.h:2070-2078).
__global__ void kb_prim_sample_kernel(const float* __restrict__ table, int size, int nargs, float phase, int n, float* __restrict__ out) {
	if (threadIdx.x || blockIdx.x) return;
	float position = nargs >= 2 ? phase * (float)44100 : 0.f;
	const float increment = 1.f, offset = 0.f;
	for (int s = 0; s < n; s++) {
		if (!(increment >= (float)size)) { position += increment; if (position > (float)size) position -= (float)size; }
		const float off = position + offset, fl = floorf(off), frac = off - fl;
		const int i = (int)off, j = (i == size - 1) ? 0 : i + 1;
		out[s] = table[i] * (1.f - frac) + table[j] * frac;
	}
}
// kinds (tests/cases.py): 0 Biquad::LPF 1 Biquad::HPF 2 OnePole::LPF 3 OnePole::HPF 4 Biquad::BPF 5 Biquad::BRF 6 Biquad::APF
// 7 Butterworth::LPF<1> 8 Butterworth::LPF<2> 9 DCF 10 IIR<1> 11 IIR<2>; one-pole coefficients (expf / tanf) come from the host
// 12 Modifiers::Modal, 13 / 14 Envelope::Follower peak / rms, 15 / 16 Follower::Window<64> mean / rms: coefficients hc from the host
// libm (set() is event-rate code)
__global__ void kb_prim_filter_kernel(int kind, int nset, const float* f, const float* Q, KbFs fs, int n, const float* in, float* out, float* coeffs,
                                      KbOnePole op, float4 hc, const float* __restrict__ op_sets = nullptr) {
	if (threadIdx.x || blockIdx.x) return;
	if (kind == 12) {   // Modal::input + process  klang.h:5847-5856: in *= gain; out = in + a1*y1 + a2*y2
		const float a1 = hc.x, a2 = hc.y, gain = hc.z;
		float y1 = 0.f, y2 = 0.f;
		for (int s = 0; s < n; s++) {
			const float x = in[s] * gain;
			const float o = x + a1 * y1 + a2 * y2;
			y2 = y1; y1 = o;
			out[s] = o;
		}
		coeffs[0] = a1; coeffs[1] = a2; coeffs[2] = gain; coeffs[3] = y1; coeffs[4] = y2;
		return;
	}
	if (kind == 15 || kind == 16) { kb_window_follower_run(kind == 16, hc.x, hc.y, n, in, out, coeffs); return; }   // Follower::Window<64> mean / rms
	if (kind == 13 || kind == 14) {   // Follower::peak / rms over AR::process  klang.h:5866-5896 (abs, sqrt == fabsf, sqrtf: Q4)
		const float A = hc.x, R = hc.y;
		float ar = 0.f;
		for (int s = 0; s < n; s++) {
			const float x = kind == 13 ? fabsf(in[s]) : in[s] * in[s];
			const float smoothing = x > ar ? A : R;
			ar = ar + smoothing * (x - ar);
			out[s] = kind == 13 ? ar : sqrtf(ar);
		}
		coeffs[0] = A; coeffs[1] = R; coeffs[2] = ar; coeffs[3] = 0.f; coeffs[4] = 0.f;
		return;
	}
	if (kind >= 9) {   // 9 Filters::DCF (f = r), 10 IIR<1> (f = coefficient), 11 IIR<2> (f = a1, Q = a2)   klang.h:5387-5446
		float r = 0.995f, z = 0.f, a = 1.f, b = 0.f, a2[2] = { 0.f, 0.f }, y2[2] = { 0.f, 0.f }, o = 0.f;
		for (int s = 0; s < n; s++) {
			if (s < nset) { if (kind == 9) r = f[s]; else if (kind == 10) { a = f[s]; b = 1.f - a; } else { a2[0] = f[s]; a2[1] = Q[s]; } }
			const float x = in[s];
			if (kind == 9) { o = x - z + r * o; z = x; }                       // DCF::process
			else if (kind == 10) o = x * a + o * b;                           // IIR<1>::process
			else { o = x; o -= a2[0] * y2[0]; o -= a2[1] * y2[1]; y2[1] = y2[0]; y2[0] = o; }   // IIR<2>::process
			out[s] = o;
		}
		coeffs[0] = kind == 9 ? r : kind == 10 ? a : a2[0]; coeffs[1] = kind == 9 ? z : kind == 10 ? b : a2[1];
		coeffs[2] = kind == 11 ? y2[0] : 0.f; coeffs[3] = kind == 11 ? y2[1] : 0.f; coeffs[4] = 0.f;
		return;
	}
	const int bq = kind == 0 ? KB_BQ_LPF : kind == 1 ? KB_BQ_HPF : kind == 4 ? KB_BQ_BPF : kind == 5 ? KB_BQ_BRF : kind == 6 ? KB_BQ_APF : kind == 8 ? KB_BQ_BW2 : -1;
	if (bq >= 0) {
		KbBiquad b; kb_biquad_construct(b, bq);
		for (int s = 0; s < n; s++) {
			if (s < nset) {
				if (Q) { if (bq == KB_BQ_APF) kb_apf_set(fs, b, f[s], Q[s]); else kb_biquad_set(fs, b, f[s], Q[s]); }
				else kb_biquad_set_f(fs, b, f[s]);
			}
			out[s] = kb_biquad_tick(b, in[s]);
		}
		coeffs[0] = b.b0; coeffs[1] = b.b1; coeffs[2] = b.b2; coeffs[3] = b.a1; coeffs[4] = b.a2;
	} else {
		// per-sample set(): the coefficient triples (b0, b1, a1) of every set() come from the host libm (expf / tanf), one per sample
		for (int s = 0; s < n; s++) {
			if (op_sets && s < nset) { op.b0 = op_sets[3 * s]; op.b1 = op_sets[3 * s + 1]; op.a1 = op_sets[3 * s + 2]; }
			out[s] = kb_onepole_tick(op, in[s]);
		}
		coeffs[0] = op.b0; coeffs[1] = op.b1; coeffs[2] = 0; coeffs[3] = op.a1; coeffs[4] = 0;
	}
}
__global__ void kb_prim_env_kernel(KbEnv e, KbFs fs, int n, int release_at, float rt, float rl, int adsr, float* out, int* stage) {
	if (threadIdx.x || blockIdx.x) return;
	// ticks are produced by the run-length form (kb_envr_run) in ragged chunks; the stage is sampled per tick by
	// running chunks of one sample around the positions a second pass needs
	KbEnv e2 = e;
	for (int s = 0; s < n; s++) {
		if (s == release_at) { if (adsr) kb_adsr_release(fs, e); else kb_env_release(fs, e, rt, rl); }
		out[s] = kb_env_tick(fs, e);
		stage[s] = e.stage;
	}
	int s = 0, chunk = 1;
	while (s < n) {
		int len = min(chunk, n - s);
		if (release_at >= s && release_at < s + len && release_at != s) len = release_at - s;
		if (s == release_at) { if (adsr) kb_adsr_release(fs, e2); else kb_env_release(fs, e2, rt, rl); }
		KbEnvR r; kb_envr_load(r, e2);
		float tmp[64];
		kb_envr_run(fs, r, e2.px, e2.py, tmp, len);
		kb_envr_store(r, e2);
		// any disagreement with the per-tick form is made visible as a NaN
		for (int i = 0; i < len; i++) if (__float_as_uint(tmp[i]) != __float_as_uint(out[s + i])) out[s + i] = __int_as_float(0x7fc00000);
		s += len;
		chunk = chunk % 61 + 3;
	}
}
__global__ void kb_prim_math_kernel(int fn, int n, const float* x, float* out) {
	for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
		out[i] = fn == 0 ? kb_sinf(x[i]) : fn == 1 ? kb_cosf(x[i]) : fn == 2 ? kb_tanhf(x[i]) : kb_expf(x[i]);
}

// Stereo::Delay<1000> (klang.h:4647-4700), sample by sample: `x >> delay` writes both lines at the same position (Bank<Delay,2>,
// klang.h:2901-2926), then tap(float df[s]) reads both channels at the LEFT line's position with the (1 - f) / f form.
__global__ void kb_prim_stereo_delay_kernel(int n, const float* inl, const float* inr, const float* df, float* ringl, float* ringr, float* outl, float* outr) {
	if (threadIdx.x || blockIdx.x) return;
	KbDelay l, r; kb_delay_construct(l, 1000, 0); kb_delay_construct(r, 1000, 0);
	for (int s = 0; s <= 1000; s++) { ringl[s] = 0.f; ringr[s] = 0.f; }
	for (int s = 0; s < n; s++) {
		kb_delay_write(l, ringl, inl[s]); kb_delay_write(r, ringr, inr[s]);
		kb_sdelay_tap_f(l, ringl, ringr, df[s], outl[s], outr[s]);
	}
}
// Control::set then Control::smooth per sample (klang.h:1715-1728), from a Dial(lo, hi, initial) (klang.h:1796-1799: smoothed starts at 0)
__global__ void kb_prim_control_smooth_kernel(float lo, float hi, float initial, int n, const float* values, float* out) {
	if (threadIdx.x || blockIdx.x) return;
	KbControl c = { lo, hi, initial, 0.f };
	for (int s = 0; s < n; s++) { kb_control_set(c, values[s]); out[s] = kb_control_smooth(c); }
}
// Envelope::at(t) (klang.h:3929-3942): static breakpoint lookup, thread = query
__global__ void kb_prim_envelope_at_kernel(int npts, const float* xy, int n, const float* t, float* out) {
	__shared__ float px[KB_ENV_MAXPTS], py[KB_ENV_MAXPTS];
	if (threadIdx.x < npts) { px[threadIdx.x] = xy[2 * threadIdx.x]; py[threadIdx.x] = xy[2 * threadIdx.x + 1]; }
	__syncthreads();
	for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) out[i] = kb_env_at(px, py, npts, t[i]);
}
// Wavetable (klang.h:3627-3676): the 2048-entry table the device oscillators of kb_prim_osc_kernel read, entry by entry (buffer::operator[](int),
// klang.h:2060-2068) — the table is filled on the host (Wavetable::operator=(Oscillator) is constructor code) and lives in HBM
__global__ void kb_prim_wavetable_kernel(const float* table, float* out) {
	for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < 2048; i += gridDim.x * blockDim.x) out[i] = table[i];
}

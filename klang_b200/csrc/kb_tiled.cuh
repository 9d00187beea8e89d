// klang-b200 — pipelined synth-voice kernels.
//
// A voice is a few short recurrences (envelope ramps, the filter) wrapped around work that is a pure function of
// the sample index (band-limited oscillators on an integer phase ramp, filter coefficients = f(cutoff)).  With
// only ~1000 voices a lane-per-voice loop leaves the chip idle and is bound by the dependent-issue latency of
// the whole per-sample expression.  Here a CTA owns G voices, the block is cut into tiles of T samples,
// and four stages work on four consecutive tiles at once, in lock step (one __syncthreads per tick):
//   A  recurrences that feed others   warps 0-1, lane = voice   tile k     (Envelope::process, klang.h:4018-4051;
//                                                                           warp 0 the modulation envelopes, warp 1 the ADSRs)
//   B  time-parallel work             warps 3.., thread=(v,t)   tile k-1   (OSM samples, Biquad::set / TB303 coefficients)
//   C  the filter recurrence          warp 2, lane = voice      tile k-2   (only the 2- / 5-state update is serial)
//   D  output                         warps 3.., thread=(v,t)   tile k-3   (soft clip, *= adsr, coalesced stores)
// so a tick costs max(A, B+D, C) instead of their sum, and the serial stages hide behind the parallel one.
// Arithmetic per sample is the reference's, operation for operation (bit-exact vs the oracle); only the order in
// which independent samples are computed changes.  Shared rows are padded to T+1 floats so that the lane=voice
// stages (stride T+1) and the thread=(voice,t) stages (stride 1) are both bank-conflict free; stage hand-over
// buffers are double buffered (A->B, B->C, C->D).
#pragma once
#include "kb_graphs.cuh"
#include "kb_sync.cuh"

#define KB_TILE_T 128      // samples per tile
// G = voices per CTA and NT = threads per CTA are template parameters: the serial stages cost the same for any G <= 32
// (one lane per voice), so G is chosen as large as the voice count allows while still filling the SMs.

template <int G> struct KbTileRows { float r[G][KB_TILE_T + 1]; };

template <int G> struct KbTileCommon {
	float px[2 * G][KB_ENV_MAXPTS], py[2 * G][KB_ENV_MAXPTS];   // envelope breakpoints of the A lanes
	int active[G];
};

// prologue shared by the kernels: active flags, and the envelope lanes (warp 0: lanes 0..G-1 first envelope,
// lanes G..2G-1 the ADSR) load their scalar state into registers and their breakpoints into shared memory
template <int G, class VOICE>
KB_D void kb_tile_prologue(KbTileCommon<G>& c, VOICE* voices, KbVoiceHdr* hdr, int v0, int total) {
	if (threadIdx.x < G) {
		const int v = v0 + threadIdx.x;
		const int act = (v < total && hdr[v].stage != KB_NOTE_OFF) ? 1 : 0;
		c.active[threadIdx.x] = act;
		if (v < total) hdr[v].active = act;
	}
	__syncthreads();
}
template <int G>
KB_D void kb_tile_load_env(KbTileCommon<G>& c, int slot, const KbEnv& src, KbEnvR& e) {
	kb_envr_load(e, src);
	for (int p = 0; p < KB_ENV_MAXPTS; p++) { c.px[slot][p] = src.px[p]; c.py[slot][p] = src.py[p]; }
}

// ----------------------------------------------------------------------------------- Subtractive / Filter.k
// Filter.k:29-36.  A: env (cutoff) and adsr.  B: Biquad::Filter::set(cutoff, 10) + LPF::init (klang.h:5584-5600,
// 5658-5665) and the oscillator sample.  C: Filter::process (klang.h:5605-5612) and `out *= adsr`.  D: stores.
// Biquad::set is a pure function of (f, Q) and is re-evaluated for every sample; the reference's "unchanged (f,Q)"
// early-out returns the same coefficients (Q is the constant 10, so the first set after reset() always computes).
template <int G> struct alignas(16) KbTileRowsA { float r[G][KB_TILE_T + 4]; };     // row stride 132 floats: lane=voice 128-bit stores hit banks 4v..4v+3
template <int G> struct KbTileRows4 { float4 r[G][KB_TILE_T + 1]; };   // lane=voice reads 16 B from banks 4v..4v+3: conflict-free per quarter warp
template <int G> struct KbTileRows2 { float2 r[G][KB_TILE_T + 1]; };
template <int G> struct KbSubSmem {
	KbTileCommon<G> c;
	KbTileRows4<G> coef[2];          // B -> C: (b0*in, b1*in, a1, a2) of every sample: one 128-bit load per step is all C reads
	KbTileRows<G> cut[2], amp[4];    // A -> B cutoff; A -> D adsr level (three ticks later)
	KbTileRowsA<G> out[2];           // C -> D filter output (16-byte aligned rows: C stores four samples at a time)
	float4 lastc[G];                 // (b0, b1, a1, a2) of the block's last sample, for the state write-back
	KbOsm osc[G];
};
// LAYOUT 0: the serial roles are warps 0, 1, 2 (one per SM sub-partition, each sharing its issue slots with worker warps).
// LAYOUT 2 (NT = 768): the filter recurrence (C, warp 0) has sub-partition 0 (warp id mod 4) to itself — warps 4, 8, .. stay
//           idle: its 16-cycle dependent chain is what bounds the kernel and every issue slot it loses to another warp
//           stretches it (measured: ~20 cycles per sample alone, ~40 when sharing).  Warp 1 runs BOTH envelopes (cutoff
//           envelope on lanes 0..G-1, ADSR on lanes G..2G-1: the run loop is one instruction stream for every envelope mode)
//           and shares sub-partition 1 with four worker warps; sub-partitions 2 and 3 hold six worker warps each: 16 worker
//           warps = 512 threads = exactly two rounds over an 8-voice x 128-sample tile.
template <int LAYOUT> KB_D int kb_tile_role(int warp) {
	if (LAYOUT == 0) return warp < 3 ? warp : -1;
	return warp == 0 ? 2 : warp == 1 ? 0 : -1;
}
template <int LAYOUT> KB_D bool kb_tile_is_worker(int warp) {
	if (LAYOUT == 0) return warp >= 3;
	return (warp & 3) >= 2 || ((warp & 3) == 1 && warp >= 9);
}
template <int LAYOUT, int NT> KB_D int kb_tile_worker_tid(int warp, int lane) {
	if (LAYOUT == 0) return (warp - 3) * 32 + lane;
	constexpr int R = NT / 128;                                       // warps per sub-partition
	const int sp = warp & 3, row = warp >> 2;
	return (sp == 1 ? row - 2 : sp == 2 ? (R - 2) + row : (2 * R - 2) + row) * 32 + lane;
}
template <int LAYOUT, int NT> constexpr int kb_tile_worker_threads_v = LAYOUT == 0 ? NT - 96 : (3 * (NT / 128) - 2) * 32;
template <int G, int NT, int LAYOUT = 0>
__global__ void __launch_bounds__(NT, 1) kb_sub_tiled_kernel(KbSubVoice* __restrict__ voices, KbVoiceHdr* __restrict__ hdr,
                                                                        float* __restrict__ dst, int n, int total, KbFs fs, long long* __restrict__ trace = nullptr) {
	constexpr int T = KB_TILE_T;
	extern __shared__ __align__(16) unsigned char kb_smem[];
	KbSubSmem<G>& S = *reinterpret_cast<KbSubSmem<G>*>(kb_smem);
	const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
	const int v0 = blockIdx.x * G;
	kb_tile_prologue(S.c, voices, hdr, v0, total);

	const int role = kb_tile_role<LAYOUT>(warp);                      // 0 = A (cutoff envelope), 1 = A (ADSR), 2 = C (filter), -1 = worker / idle
	constexpr bool MERGED_A = LAYOUT == 2 && 2 * G <= 32;             // role 0 runs both envelopes of a voice on lanes v and G + v
	const int a_sub = (MERGED_A && role == 0 && lane >= G) ? 1 : 0;
	const int role_voice = lane - a_sub * G;
	const bool role_ok = role_voice < G && S.c.active[role_voice < G ? role_voice : 0];
	const bool worker = kb_tile_is_worker<LAYOUT>(warp);
	const bool first_worker = worker && kb_tile_worker_tid<LAYOUT, NT>(warp, lane) < 32;
	const bool is_env = role == 0 && a_sub == 0 && role_ok, is_adsr = (role == 1 || a_sub == 1) && role_ok, is_flt = role == 2 && role_ok;
	const int slot = (is_adsr ? G : 0) + role_voice;                  // breakpoint slot of the A lanes
	KbEnvR env;
	float z0 = 0.f, z1 = 0.f;
	if (is_env) kb_tile_load_env(S.c, slot, voices[v0 + role_voice].env, env);
	if (is_adsr) kb_tile_load_env(S.c, slot, voices[v0 + role_voice].adsr, env);
	if (is_flt) { const KbBiquad& b = voices[v0 + role_voice].filter; z0 = b.z0; z1 = b.z1; }
	if (first_worker && lane < G && S.c.active[lane]) S.osc[lane] = voices[v0 + lane].osc;
	__syncthreads();

	const int ntiles = (n + T - 1) / T;
	const int wtid = kb_tile_worker_tid<LAYOUT, NT>(warp, lane);         // the B/D worker threads
	constexpr int wthreads = kb_tile_worker_threads_v<LAYOUT, NT>;
	static_assert(LAYOUT != 2 || NT % 128 == 0, "layout 2 needs whole rows of four warps");
	// measurement aid (KB_C2_TRACE=<file>, tools/c2_trace.py): CTA 0 stamps clock64() at the start of every tick and when each role has finished its
	// part of it: trace[(row * 64 + tick) * 2 + {0 start, 1 end}], rows 0 = A, 1 = C, 2 = a worker after B, 3 = the same worker after D
	#define KB_C2_TR(row_, k_, ph_) do { if (trace && blockIdx.x == 0 && (k_) < 64) trace[(((row_) * 64 + (k_)) * 2 + (ph_))] = clock64(); } while (0)
	const bool tr_a = (role == 0 || role == 1) && lane == 0, tr_c = role == 2 && lane == 0, tr_w = first_worker && lane == 0;
	for (int k = 0; k < ntiles + 3; k++) {
		if (tr_a) KB_C2_TR(0, k, 0);
		if (tr_c) KB_C2_TR(1, k, 0);
		if (tr_w) KB_C2_TR(2, k, 0);
		if (role == 0 || role == 1) {                                    // ---- A, tile k
			if ((is_env || is_adsr) && k < ntiles) {
				const int steps = min(T, n - k * T);
				float* row = is_env ? S.cut[k & 1].r[role_voice] : S.amp[k & 3].r[role_voice];
				kb_envr_run(fs, env, S.c.px[slot], S.c.py[slot], row, steps);
			}
		} else if (role == 2) {                                          // ---- C, tile k-2
			const int c = k - 2;
			if (is_flt && c >= 0 && c < ntiles) {
				const int steps = min(T, n - c * T), v = role_voice;
				const float4* pc = S.coef[c & 1].r[v];
				float* po = S.out[c & 1].r[v];
				// Filter::process (klang.h:5605-5612) with the input products b0*in, b1*in (b2 = b0) formed by B: per step one 128-bit
				// load, the four dependent operations of the recurrence, two independent ones and a store.  Groups of 4 steps; the
				// operands of the NEXT group are loaded before the current group's updates, so no shared-memory latency sits on
				// the recurrence (reads past `steps` stay inside the shared-memory struct and are unused)
				// two register sets used alternately (no copies): while set A is consumed set B is loaded, and vice versa
				auto group = [&](const float4 (&cf)[4], int t) {
					float y[4];
					#pragma unroll
					for (int j = 0; j < 4; j++) {
						y[j] = cf[j].x + z0;                                 // y = b0*in + z0
						z0 = cf[j].y - cf[j].z * y[j] + z1;                  // z0 = b1*in - a1*y + z1
						z1 = cf[j].x - cf[j].w * y[j];                       // z1 = b2*in - a2*y   (LPF: b2 == b0)
					}
					*reinterpret_cast<float4*>(po + t) = make_float4(y[0], y[1], y[2], y[3]);
				};
				float4 ca[4], cb[4];
				#pragma unroll
				for (int j = 0; j < 4; j++) ca[j] = pc[j];
				int t = 0;
				for (; t + 8 <= steps; t += 8) {
					#pragma unroll
					for (int j = 0; j < 4; j++) cb[j] = pc[t + 4 + j];
					group(ca, t);
					#pragma unroll
					for (int j = 0; j < 4; j++) ca[j] = pc[t + 8 + j];
					group(cb, t + 4);
				}
				for (; t < steps; t++) {                                 // ragged tail (only the last tile of a block)
					const float4 c1 = pc[t];
					const float y = c1.x + z0;
					z0 = c1.y - c1.z * y + z1;
					z1 = c1.x - c1.w * y;
					po[t] = y;
				}
			}
		} else if (worker) {
			const int b = k - 1, d = k - 3;
			if (b >= 0 && b < ntiles) {                                      // ---- B, tile k-1
				const int steps = min(T, n - b * T);
				for (int item = wtid; item < G * T; item += wthreads) {
					const int v = item / T, t = item % T;
					if (t < steps && S.c.active[v]) {
						const float f = S.cut[b & 1].r[v][t];
						const float w = f * fs.w;
						float sin0, cos0;
						kb_sincosf(w, sin0, cos0);
						const float a = sin0 / (2.f * 10.f);
						const float inv = kb_const_inv(1.f + a);
						float4 cf;
						cf.z = inv * (-2.f * cos0);
						cf.w = inv * (1.f - a);
						cf.x = inv * (1.f - cos0) * 0.5f;
						cf.y = inv * (1.f - cos0);
						if (b * T + t == n - 1) S.lastc[v] = cf;
						const float in = kb_osm_at(S.osc[v], (uint32_t)(b * T + t));
						cf.x = cf.x * in; cf.y = cf.y * in;                  // the input products of Filter::process
						S.coef[b & 1].r[v][t] = cf;
					}
				}
			}
			if (tr_w) KB_C2_TR(2, k, 1);
			if (d >= 0 && d < ntiles) {                                      // ---- D, tile k-3
				const int steps = min(T, n - d * T);
				for (int item = wtid; item < G * T; item += wthreads) {
					const int v = item / T, t = item % T;
					if (t < steps && v0 + v < total)                             // out *= adsr++   Filter.k:33
						dst[(size_t)(v0 + v) * n + d * T + t] = S.c.active[v] ? S.out[d & 1].r[v][t] * S.amp[d & 3].r[v][t] : 0.f;
				}
			}
		}
		if (tr_a) KB_C2_TR(0, k, 1);
		if (tr_c) KB_C2_TR(1, k, 1);
		if (tr_w) KB_C2_TR(3, k, 1);
		__syncthreads();
	}
	#undef KB_C2_TR

	// write the state back
	if (is_env) { kb_envr_store(env, voices[v0 + role_voice].env); voices[v0 + role_voice].filter.f = env.out; voices[v0 + role_voice].filter.Q = 10.f; }
	if (is_adsr) {
		kb_envr_store(env, voices[v0 + role_voice].adsr);
		if (env.stage == KB_ENV_OFF) hdr[v0 + role_voice].stage = KB_NOTE_OFF;       // if (adsr.finished()) stop()   Filter.k:34-35
	}
	if (is_flt) {
		KbBiquad& b = voices[v0 + role_voice].filter;
		const float4 lc = S.lastc[role_voice];
		b.z0 = z0; b.z1 = z1; b.b0 = lc.x; b.b2 = lc.x; b.b1 = lc.y; b.a1 = lc.z; b.a2 = lc.w;
	}
	if (first_worker && lane < G && S.c.active[lane]) {
		KbOsm o = S.osc[lane];
		kb_osm_advance(o, (uint32_t)n);
		voices[v0 + lane].osc.offset = o.offset;
		voices[v0 + lane].osc.state = o.state;
	}
}

// The event upload folded into the voice kernel (round 2: no copy-engine operation and no scatter launch in front of the voice kernel — a
// kernel queued behind a host-to-device copy starts several microseconds after it).  Voices re-written by the host since the last block wait
// packed in PINNED HOST memory, [count][hdr | blob]; their voice indices travel in the kernel's parameters.  A CTA whose voices are listed
// pulls their records over PCIe into shared memory (one round trip, and only for the CTAs concerned), LOADS those voices from there, and
// copies the records to the voices' slots by fire-and-forget stores (the slot is next read by the following launch; this kernel's own state
// write-back comes many barriers later).  The header's `active` word is written by the prologue, not by the copy.  More than KB_STAGED_MAX
// re-written voices, and every other kernel, take the copy + kb_scatter_voices_kernel path.
#define KB_STAGED_MAX 128
struct KbStaged { const unsigned char* records; int count, voice_bytes; int index[KB_STAGED_MAX]; };
template <int G, class VOICE> struct KbStagedSmem { int src[G]; unsigned rec[G][(sizeof(KbVoiceHdr) + sizeof(VOICE)) / 4]; };
template <int G, class VOICE>
KB_D void kb_tile_scatter(const KbStaged& sg, KbVoiceHdr* hdr, VOICE* voices, int v0, KbStagedSmem<G, VOICE>& m) {
	if (threadIdx.x < G) m.src[threadIdx.x] = -1;
	if (sg.count <= 0) return;                                           // (uniform over the grid)
	__syncthreads();
	for (int k = threadIdx.x; k < sg.count; k += blockDim.x) {
		const int v = sg.index[k];
		if (v >= v0 && v < v0 + G) m.src[v - v0] = k;                     // (a voice is listed once)
	}
	__syncthreads();
	constexpr int rec_words = (int)(sizeof(KbVoiceHdr) + sizeof(VOICE)) / 4, hdr_words = (int)sizeof(KbVoiceHdr) / 4;
	for (int i = threadIdx.x; i < G * rec_words; i += blockDim.x) {
		const int g = i / rec_words, w = i % rec_words, k = m.src[g];
		if (k >= 0) m.rec[g][w] = reinterpret_cast<const unsigned*>(sg.records)[(size_t)k * rec_words + w];
	}
	__syncthreads();
	for (int i = threadIdx.x; i < G * rec_words; i += blockDim.x) {
		const int g = i / rec_words, w = i % rec_words;
		if (m.src[g] < 0 || w == (int)(offsetof(KbVoiceHdr, active) / 4)) continue;
		if (w < hdr_words) reinterpret_cast<unsigned*>(hdr + v0 + g)[w] = m.rec[g][w];
		else reinterpret_cast<unsigned*>(voices + v0 + g)[w - hdr_words] = m.rec[g][w];
	}
}
template <int G, class VOICE>
KB_D const KbVoiceHdr& kb_tile_hdr_src(const KbStagedSmem<G, VOICE>& m, const KbVoiceHdr* hdr, int v0, int i) {
	return m.src[i] < 0 ? hdr[v0 + i] : *reinterpret_cast<const KbVoiceHdr*>(m.rec[i]);
}
template <int G, class VOICE>
KB_D const VOICE& kb_tile_voice_src(const KbStagedSmem<G, VOICE>& m, const VOICE* voices, int v0, int i) {
	return m.src[i] < 0 ? voices[v0 + i] : *reinterpret_cast<const VOICE*>(m.rec[i] + sizeof(KbVoiceHdr) / 4);
}

// One round trip instead of three (kb_sub_mbar_kernel): every record of the CTA's voices — header and state, from the pinned staging slot for a
// re-written voice, else from its slot in HBM — is requested in ONE cooperative pass into shared memory (the PCIe read and the HBM reads overlap);
// all later prologue reads (note stage, envelopes, filter, oscillator) come from there.  The staged records are written through to their slots as
// in kb_tile_scatter.  `m.src[g]` ends >= 0 for every voice that exists, so kb_tile_hdr_src / kb_tile_voice_src always pick the shared copy.
template <int G, class VOICE>
KB_D void kb_tile_gather(const KbStaged& sg, KbVoiceHdr* hdr, VOICE* voices, int v0, int total, KbStagedSmem<G, VOICE>& m, int* from_host) {
	if (threadIdx.x < G) { m.src[threadIdx.x] = -1; from_host[threadIdx.x] = 0; }
	__syncthreads();
	for (int k = threadIdx.x; k < sg.count; k += blockDim.x) {
		const int v = sg.index[k];
		if (v >= v0 && v < v0 + G) { m.src[v - v0] = k; from_host[v - v0] = 1; }
	}
	__syncthreads();
	constexpr int rec_words = (int)(sizeof(KbVoiceHdr) + sizeof(VOICE)) / 4, hdr_words = (int)sizeof(KbVoiceHdr) / 4;
	for (int i = threadIdx.x; i < G * rec_words; i += blockDim.x) {
		const int g = i / rec_words, w = i % rec_words, k = m.src[g];
		if (v0 + g >= total) continue;
		if (k >= 0) m.rec[g][w] = reinterpret_cast<const unsigned*>(sg.records)[(size_t)k * rec_words + w];
		else m.rec[g][w] = w < hdr_words ? reinterpret_cast<const unsigned*>(hdr + v0 + g)[w] : reinterpret_cast<const unsigned*>(voices + v0 + g)[w - hdr_words];
	}
	__syncthreads();
	for (int i = threadIdx.x; i < G * rec_words; i += blockDim.x) {
		const int g = i / rec_words, w = i % rec_words;
		if (!from_host[g] || w == (int)(offsetof(KbVoiceHdr, active) / 4)) continue;
		if (w < hdr_words) reinterpret_cast<unsigned*>(hdr + v0 + g)[w] = m.rec[g][w];
		else reinterpret_cast<unsigned*>(voices + v0 + g)[w - hdr_words] = m.rec[g][w];
	}
	if (threadIdx.x < G) m.src[threadIdx.x] = (v0 + (int)threadIdx.x < total) ? 0 : -1;   // (src is read again only behind the caller's next barrier)
}

// ---- the same four stages WITHOUT the lock step (round 2): every role runs its own tile loop and meets the others through progress
// counters in shared memory (kb_sync.cuh: st.release by one thread behind the role's own barrier, ld.acquire polling), over hand-over
// buffers four tiles deep (A -> D: eight), so a role waits only when the role it depends on is really behind — the tick of the lock-step
// kernel was max(A, B + D, C) PLUS the skew of one CTA-wide barrier per tile.  768 threads: warp 0 = C alone on SM sub-partition 0 (with
// A_SP0 the envelope warp is warp 4, the only other warp there — two latency-bound chains interleave in the issue slots neither fills;
// without it A is warp 1 and shares sub-partition 1 with worker warps), 2 * G worker warps on sub-partitions 1-3 = exactly two rounds
// over a G x 128 tile (G = 7: 147 CTAs for 1024 voices, one per SM, and an eighth less B + D work per CTA than G = 8 on 128 CTAs).
// Arithmetic and its order per voice are those of kb_sub_tiled_kernel: bit-identical output and state.
template <int G> struct KbSubFlowSmem {
	KbTileCommon<G> c;
	KbTileRows4<G> coef[4];          // B -> C
	KbTileRowsA<G> cut[4], amp[8];   // A -> B, A -> D (16-byte aligned rows: the envelope warp stores 128-bit words)
	KbTileRowsA<G> out[4];           // C -> D
	float4 lastc[G];
	KbOsm osc[G];
	int a_done, b_done, c_done, d_done;   // tiles finished by each stage
	KbStagedSmem<G, KbSubVoice> sc;       // the staged upload of this CTA's voices (kb_tile_scatter)
};
template <int G, bool A_SP0>
__global__ void __launch_bounds__(768, 1) kb_sub_flow_kernel(KbSubVoice* __restrict__ voices, KbVoiceHdr* __restrict__ hdr,
                                                             float* __restrict__ dst, int n, int total, KbFs fs, const __grid_constant__ KbStaged staged, long long* __restrict__ trace = nullptr, int variant = 0) {
	const bool sig_relaxed = (variant & 1) != 0;
	constexpr int T = KB_TILE_T, W = 2 * G, wthreads = W * 32, LAG = 3;
	static_assert(2 * G <= 32 && W <= 17, "both envelopes of the G voices in one warp; worker warps on three sub-partitions of six rows");
	extern __shared__ __align__(16) unsigned char kb_smem[];
	KbSubFlowSmem<G>& S = *reinterpret_cast<KbSubFlowSmem<G>*>(kb_smem);
	const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
	const int v0 = blockIdx.x * G;
	const long long t_entry = trace ? clock64() : 0;
	if (tid == 0) { S.a_done = 0; S.b_done = 0; S.c_done = 0; S.d_done = 0; }
	kb_tile_scatter(staged, hdr, voices, v0, S.sc);
	if (tid < G) {                                                       // kb_tile_prologue with the staged headers
		const int v = v0 + tid;
		const int act = (v < total && kb_tile_hdr_src(S.sc, hdr, v0, tid).stage != KB_NOTE_OFF) ? 1 : 0;
		S.c.active[tid] = act;
		if (v < total) hdr[v].active = act;
	}
	__syncthreads();

	const int sp = warp & 3, row = warp >> 2;
	const bool is_c_warp = warp == 0, is_a_warp = A_SP0 ? warp == 4 : warp == 1;
	const int widx = sp == 0 ? -1 : (A_SP0 ? row * 3 + (sp - 1) : row * 3 + (sp - 1) - 1);      // worker warps, spread evenly over sub-partitions 1-3
	const bool worker = widx >= 0 && widx < W && !is_a_warp;
	const int wtid = widx * 32 + lane;
	const int a_sub = (is_a_warp && lane >= G) ? 1 : 0;
	const int role_voice = lane - a_sub * G;
	const bool role_ok = role_voice < G && S.c.active[role_voice < G ? role_voice : 0];
	const bool is_env = is_a_warp && a_sub == 0 && role_ok, is_adsr = is_a_warp && a_sub == 1 && role_ok, is_flt = is_c_warp && role_ok;
	const int slot = (is_adsr ? G : 0) + role_voice;
	KbEnvR env;
	float z0 = 0.f, z1 = 0.f;
	if (is_env) kb_tile_load_env(S.c, slot, kb_tile_voice_src(S.sc, voices, v0, role_voice).env, env);
	if (is_adsr) kb_tile_load_env(S.c, slot, kb_tile_voice_src(S.sc, voices, v0, role_voice).adsr, env);
	if (is_flt) { const KbBiquad& b = kb_tile_voice_src(S.sc, voices, v0, role_voice).filter; z0 = b.z0; z1 = b.z1; }
	if (worker && wtid < G && S.c.active[wtid]) S.osc[wtid] = kb_tile_voice_src(S.sc, voices, v0, wtid).osc;
	__syncthreads();

	const int ntiles = (n + T - 1) / T;
	const long long t_loop = trace ? clock64() : 0;
	#define KB_C2_TR(row_, k_, ph_) do { if (trace && blockIdx.x == 0 && (k_) < 64) trace[(((row_) * 64 + (k_)) * 2 + (ph_))] = clock64(); } while (0)
	if (is_a_warp) {                                                     // ---- A: both envelopes of every voice, tile after tile
		for (int k = 0; k < ntiles; k++) {
			if (k >= 4) kb_wait_ge(&S.b_done, k - 3);                    // cut[k & 3] was tile k-4's
			if (k >= 8) kb_wait_ge(&S.d_done, k - 7);                    // amp[k & 7] was tile k-8's
			if (lane == 0) KB_C2_TR(0, k, 0);
			if (is_env || is_adsr) {
				const int steps = min(T, n - k * T);
				float* rowp = is_env ? S.cut[k & 3].r[role_voice] : S.amp[k & 7].r[role_voice];
				if (variant & 32) kb_envr_run(fs, env, S.c.px[slot], S.c.py[slot], rowp, steps);
				else kb_envr_run16<true>(fs, env, S.c.px[slot], S.c.py[slot], rowp, steps);
			}
			if (lane == 0) KB_C2_TR(8, k, 0);
			__syncwarp();
			if (lane == 0) { KB_C2_TR(8, k, 1); kb_signal_v(sig_relaxed, &S.a_done, k + 1); KB_C2_TR(0, k, 1); }
		}
	} else if (is_c_warp) {                                              // ---- C: the filter recurrence
		for (int c = 0; c < ntiles; c++) {
			kb_wait_ge(&S.b_done, c + 1);
			if (c >= 4) kb_wait_ge(&S.d_done, c - 3);                    // out[c & 3] was tile c-4's
			if (lane == 0) KB_C2_TR(1, c, 0);
			if (is_flt) {
				const int steps = min(T, n - c * T), v = role_voice;
				const float4* pc = S.coef[c & 3].r[v];
				float* po = S.out[c & 3].r[v];
				auto group = [&](const float4 (&cf)[4], int t) {         // Filter::process, klang.h:5605-5612 (see kb_sub_tiled_kernel)
					float y[4];
					#pragma unroll
					for (int j = 0; j < 4; j++) {
						y[j] = cf[j].x + z0;
						z0 = cf[j].y - cf[j].z * y[j] + z1;
						z1 = cf[j].x - cf[j].w * y[j];
					}
					*reinterpret_cast<float4*>(po + t) = make_float4(y[0], y[1], y[2], y[3]);
				};
				float4 ca[4], cb[4];
				#pragma unroll
				for (int j = 0; j < 4; j++) ca[j] = pc[j];
				int t = 0;
				for (; t + 8 <= steps; t += 8) {
					#pragma unroll
					for (int j = 0; j < 4; j++) cb[j] = pc[t + 4 + j];
					group(ca, t);
					#pragma unroll
					for (int j = 0; j < 4; j++) ca[j] = pc[t + 8 + j];
					group(cb, t + 4);
				}
				for (; t < steps; t++) {
					const float4 c1 = pc[t];
					const float y = c1.x + z0;
					z0 = c1.y - c1.z * y + z1;
					z1 = c1.x - c1.w * y;
					po[t] = y;
				}
			}
			__syncwarp();
			if (lane == 0) { kb_signal_v(sig_relaxed, &S.c_done, c + 1); KB_C2_TR(1, c, 1); }
		}
	} else if (worker) {                                                 // ---- B (tile j) then D (tile j - LAG), two rounds each
		const bool first = widx == 0;
		for (int j = 0; j < ntiles + LAG; j++) {
			// one poll per iteration by the first worker warp, the others park at the role's barrier: B needs A's tile j and its coef buffer
			// back from C (tile j-4), D needs C's tile j-LAG.  The barrier also closes the previous iteration's D for every worker.
			if (wtid == 0) KB_C2_TR(9, j, 0);
			if (first) {
				if (j < ntiles) kb_wait_ge(&S.a_done, j + 1);
				kb_wait_ge(&S.c_done, min(max(j - LAG + 1, 0), ntiles));
			}
			kb_bar_group(1, wthreads);
			if (wtid == 0) { if (j - 1 - LAG >= 0) kb_signal_v(sig_relaxed, &S.d_done, j - LAG); KB_C2_TR(2, j, 0); }
			if (j < ntiles) {
				const int b = j, steps = min(T, n - b * T);
				#pragma unroll
				for (int r = 0; r < 2; r++) {
					const int item = wtid + r * wthreads, v = item / T, t = item % T;
					if (t < steps && S.c.active[v]) {
						const float f = S.cut[b & 3].r[v][t];
						const float w = f * fs.w;
						float sin0, cos0;
						kb_sincosf(w, sin0, cos0);
						const float a = sin0 / (2.f * 10.f);
						const float inv = kb_const_inv(1.f + a);
						float4 cf;
						cf.z = inv * (-2.f * cos0);
						cf.w = inv * (1.f - a);
						cf.x = inv * (1.f - cos0) * 0.5f;
						cf.y = inv * (1.f - cos0);
						if (b * T + t == n - 1) S.lastc[v] = cf;
						const float in = kb_osm_at(S.osc[v], (uint32_t)(b * T + t));
						cf.x = cf.x * in; cf.y = cf.y * in;
						S.coef[b & 3].r[v][t] = cf;
					}
				}
				if (wtid == 0) KB_C2_TR(9, j, 1);
				kb_bar_group(2, wthreads);
				if (wtid == 0) { kb_signal_v(sig_relaxed, &S.b_done, b + 1); KB_C2_TR(2, j, 1); }
			}
			const int d = j - LAG;
			if (d >= 0 && d < ntiles) {
				const int steps = min(T, n - d * T);
				#pragma unroll
				for (int r = 0; r < 2; r++) {
					const int item = wtid + r * wthreads, v = item / T, t = item % T;
					if (t < steps && v0 + v < total)                             // out *= adsr++   Filter.k:33
						dst[(size_t)(v0 + v) * n + d * T + t] = S.c.active[v] ? S.out[d & 3].r[v][t] * S.amp[d & 7].r[v][t] : 0.f;
				}
			}
			if (wtid == 0) KB_C2_TR(3, j, 1);
		}
	}
	#undef KB_C2_TR
	__syncthreads();
	// rows 4-7 of the trace: per CTA (prologue, tile loop) cycles
	if (trace && tid == 0 && blockIdx.x < 256) { const long long t1 = clock64(); trace[(4 * 64 + blockIdx.x) * 2] = t_loop - t_entry; trace[(4 * 64 + blockIdx.x) * 2 + 1] = t1 - t_loop; }

	// write the state back
	if (is_env) { kb_envr_store(env, voices[v0 + role_voice].env); voices[v0 + role_voice].filter.f = env.out; voices[v0 + role_voice].filter.Q = 10.f; }
	if (is_adsr) {
		kb_envr_store(env, voices[v0 + role_voice].adsr);
		if (env.stage == KB_ENV_OFF) hdr[v0 + role_voice].stage = KB_NOTE_OFF;
	}
	if (is_flt) {
		KbBiquad& b = voices[v0 + role_voice].filter;
		const float4 lc = S.lastc[role_voice];
		b.z0 = z0; b.z1 = z1; b.b0 = lc.x; b.b2 = lc.x; b.b1 = lc.y; b.a1 = lc.z; b.a2 = lc.w;
	}
	if (worker && wtid < G && S.c.active[wtid]) {
		KbOsm o = S.osc[wtid];
		kb_osm_advance(o, (uint32_t)n);
		voices[v0 + wtid].osc.offset = o.offset;
		voices[v0 + wtid].osc.state = o.state;
	}
}

// ---- the decoupled stages handed over through MBARRIERS (round 2, second step; SASS SYNCS).  kb_sub_flow_kernel's progress counters are
// polled (ld.acquire + nanosleep) and its 14 worker warps meet at two named barriers per tile: ~500 of the workers' ~3150 cycles per tile
// and ~300 of the filter warp's were hand-over (clock64 stamps, profiles/r02_c2_mbar.txt).  Here a producing WARP arrives on an mbarrier
// once behind its __syncwarp() and a consumer parks in mbarrier.try_wait — no polling, and no barrier among the worker warps at all: each
// takes its own items of a tile as soon as that tile's inputs are there.
// One mbarrier ring per ROLE says "this role has finished tile n" (slot n & 3, the n-th use of a slot waits for parity (n >> 2) & 1):
//   a_done (1 arrival: the envelope warp)    b_done (W arrivals: every worker warp after its B items)    c_done (1: the filter warp)
// A worker warp's iteration j is: wait c_done(j-4) -> D(j-4) -> wait a_done(j) -> B(j) -> arrive b_done(j).  D before B, so that b_done(j)
// also says "this warp has finished D(j-4)" and no barrier for D is needed:
//   A(k) waits b_done(k-4): cut[k & 3] was read by B(k-4), amp[k & 7] by D(k-8) (before B(k-4) in every warp's order)
//   C(c) waits b_done(c):   coef[c & 3] is written, and out[c & 3] was read by D(c-4)
//   B(j) needs a_done(j) and coef[j & 3] back from C(j-4); D(d) needs C(d) and A(d) (A(d) was waited for by the same warp's B(d))
// No barrier can complete twice before a waiter has looked: the next completion of a slot needs the role that waits on it to have moved on.
// The envelope warp runs kb_envr_run_tile (one exit test per tile, 128-bit stores); the filter warp's loop is unrolled to 56 samples per
// branch and carries on across tile boundaries when the next tile's coefficients are already there (non-blocking look at b_done); the
// prologue gathers every record of the CTA's voices in one round trip (kb_tile_gather); full tiles leave the D stage two samples per
// thread.  Arithmetic and order per voice unchanged: bit-identical.
template <int G> struct KbSubMbarSmem {
	KbTileCommon<G> c;
	KbTileRows4<G> coef[4];
	KbTileRowsA<G> cut[4], amp[8];
	KbTileRowsA<G> out[4];
	float4 lastc[G];
	KbOsm osc[G];
	unsigned long long a_done[4], b_done[4], c_done[4];
	int from_host[G];
	int protocol_error;                   // CHECK instantiation: a stamp did not match (the block's output is poisoned from then on)
	KbStagedSmem<G, KbSubVoice> sc;
};
// CHECK (KB_C2_VARIANT bit 7, a second instantiation; tests/test_gpu_schedules.py): the hand-over protocol checks itself.  Every producer stamps
// the pad element of each row it is about to write with the tile number BEFORE the data; every consumer compares the stamp with the tile it
// expects before AND after it has read the row — a buffer handed over early, or overwritten while it is still being read, poisons the block's
// output with NaNs (compute-sanitizer's racecheck does not model mbarrier ordering expressed in inline PTX: it reports every buffer of the ring).
template <int G, bool CHECK = false>
__global__ void __launch_bounds__(768, 1) kb_sub_mbar_kernel(KbSubVoice* __restrict__ voices, KbVoiceHdr* __restrict__ hdr,
                                                             float* __restrict__ dst, int n, int total, KbFs fs, const __grid_constant__ KbStaged staged, long long* __restrict__ trace = nullptr, int variant = 0) {
	constexpr int T = KB_TILE_T, W = 2 * G, wthreads = W * 32, LAG = 4;
	static_assert(2 * G <= 32 && W <= 17, "both envelopes of the G voices in one warp; worker warps on three sub-partitions of six rows");
	extern __shared__ __align__(16) unsigned char kb_smem[];
	KbSubMbarSmem<G>& S = *reinterpret_cast<KbSubMbarSmem<G>*>(kb_smem);
	const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
	const int v0 = blockIdx.x * G;
	const long long t_entry = trace ? clock64() : 0;
	asm volatile("griddepcontrol.launch_dependents;");                   // the mix kernel's CTAs may take the SMs as this grid's CTAs leave (they wait for the whole grid)
	if (tid == 0) {
		S.protocol_error = 0;
		for (int i = 0; i < 4; i++) { kb_mbar_init(&S.a_done[i], 1); kb_mbar_init(&S.b_done[i], W); kb_mbar_init(&S.c_done[i], 1); }
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	kb_tile_gather(staged, hdr, voices, v0, total, S.sc, S.from_host);
	__syncthreads();
	if (tid < G) {                                                       // kb_tile_prologue from the gathered headers
		const int v = v0 + tid;
		const int act = (v < total && kb_tile_hdr_src(S.sc, hdr, v0, tid).stage != KB_NOTE_OFF) ? 1 : 0;
		S.c.active[tid] = act;
		if (v < total) hdr[v].active = act;
	}
	__syncthreads();

	// warp roles.  Sub-partition = warp & 3; the arbiter of a sub-partition prefers the HIGHER warp id.  Default: C = warp 20 and A = warp 4 on
	// sub-partition 0 (the filter chain wins the issue slot whenever it is ready, the envelopes fill the gaps), workers on sub-partitions 1-3.
	// variant bits 8-12 / 16-20 (measurement aid): C's / A's warp + 1
	const int sp = warp & 3, row = warp >> 2;
	const int c_warp = ((variant >> 8) & 31) ? ((variant >> 8) & 31) - 1 : 20, a_warp = ((variant >> 16) & 31) ? ((variant >> 16) & 31) - 1 : 4;
	const bool is_c_warp = warp == c_warp, is_a_warp = warp == a_warp;
	int widx = sp == 0 ? -1 : row * 3 + (sp - 1);                        // worker warps, spread evenly over sub-partitions 1-3
	if ((a_warp & 3) != 0) { const int aw = (a_warp >> 2) * 3 + ((a_warp & 3) - 1); if (widx == aw) widx = -1; else if (widx > aw) widx--; }
	if ((c_warp & 3) != 0) { const int cw = (c_warp >> 2) * 3 + ((c_warp & 3) - 1) - (((a_warp & 3) != 0 && a_warp < c_warp) ? 1 : 0); if (!is_a_warp) { if (widx == cw) widx = -1; else if (widx > cw) widx--; } }
	const bool worker = widx >= 0 && widx < W && !is_a_warp && !is_c_warp;
	const int wtid = widx * 32 + lane;
	const int a_sub = (is_a_warp && lane >= G) ? 1 : 0;
	const int role_voice = lane - a_sub * G;
	const bool role_ok = role_voice < G && S.c.active[role_voice < G ? role_voice : 0];
	const bool is_env = is_a_warp && a_sub == 0 && role_ok, is_adsr = is_a_warp && a_sub == 1 && role_ok, is_flt = is_c_warp && role_ok;
	const int slot = (is_adsr ? G : 0) + role_voice;
	KbEnvR env;
	float z0 = 0.f, z1 = 0.f;
	if (is_env) kb_tile_load_env(S.c, slot, kb_tile_voice_src(S.sc, voices, v0, role_voice).env, env);
	if (is_adsr) kb_tile_load_env(S.c, slot, kb_tile_voice_src(S.sc, voices, v0, role_voice).adsr, env);
	if (is_flt) { const KbBiquad& b = kb_tile_voice_src(S.sc, voices, v0, role_voice).filter; z0 = b.z0; z1 = b.z1; }
	if (worker && wtid < G && S.c.active[wtid]) S.osc[wtid] = kb_tile_voice_src(S.sc, voices, v0, wtid).osc;
	__syncthreads();

	const int ntiles = (n + T - 1) / T;
	const long long t_loop = trace ? clock64() : 0;
	#define KB_C2_TR(row_, k_, ph_) do { if (trace && blockIdx.x == 0 && (k_) < 64) trace[(((row_) * 64 + (k_)) * 2 + (ph_))] = clock64(); } while (0)
	if (is_a_warp) {                                                     // ---- A: both envelopes of every voice, tile after tile
		for (int k = 0; k < ntiles; k++) {
			if (k >= 4) kb_mbar_wait(&S.b_done[k & 3], ((k >> 2) - 1) & 1);
			if (lane == 0) KB_C2_TR(0, k, 0);
			if (is_env || is_adsr) {
				const int steps = min(T, n - k * T);
				float* rowp = is_env ? S.cut[k & 3].r[role_voice] : S.amp[k & 7].r[role_voice];
				if (CHECK) { rowp[T] = (float)k; __threadfence_block(); }
				if (variant & 64) kb_envr_run16<true>(fs, env, S.c.px[slot], S.c.py[slot], rowp, steps);
				else kb_envr_run_tile<true>(fs, env, S.c.px[slot], S.c.py[slot], rowp, steps);
			}
			if (lane == 0) KB_C2_TR(8, k, 0);
			__syncwarp();
			if (lane == 0) { kb_mbar_arrive(&S.a_done[k & 3]); KB_C2_TR(0, k, 1); }
		}
	} else if (is_c_warp) {                                              // ---- C: the filter recurrence
		// The chain runs on across tile boundaries: two groups before a tile ends the lane looks (without blocking) whether the next
		// tile's coefficients are there and, if so, loads the next tile's first group before the last group of this one — arriving and
		// waiting between two tiles cost ~300 of ~2950 cycles per tile.  A lane that did not see them waits as before.
		float4 ca[4], cb[4];
		bool have_next = false, bad = false;
		for (int c = 0; c < ntiles; c++) {
			if (!have_next) kb_mbar_wait(&S.b_done[c & 3], (c >> 2) & 1);
			if (lane == 0) KB_C2_TR(1, c, 0);
			if (is_flt) {
				const int steps = min(T, n - c * T), v = role_voice;
				const float4* pc = S.coef[c & 3].r[v];
				float* po = S.out[c & 3].r[v];
				if (CHECK) { if (pc[T].x != (float)c) bad = true; po[T] = (float)c; __threadfence_block(); }
				auto group = [&](const float4 (&cf)[4], int t) {         // Filter::process, klang.h:5605-5612 (see kb_sub_tiled_kernel)
					float y[4];
					#pragma unroll
					for (int j = 0; j < 4; j++) {
						y[j] = cf[j].x + z0;
						z0 = cf[j].y - cf[j].z * y[j] + z1;
						z1 = cf[j].x - cf[j].w * y[j];
					}
					*reinterpret_cast<float4*>(po + t) = make_float4(y[0], y[1], y[2], y[3]);
				};
				if (!have_next) {
					#pragma unroll
					for (int j = 0; j < 4; j++) ca[j] = pc[j];
				}
				have_next = false;
				int t = 0;
				if (steps == T) {
					#pragma unroll 7
					for (; t < T - 16; t += 8) {
						#pragma unroll
						for (int j = 0; j < 4; j++) cb[j] = pc[t + 4 + j];
						group(ca, t);
						#pragma unroll
						for (int j = 0; j < 4; j++) ca[j] = pc[t + 8 + j];
						group(cb, t + 4);
					}
					const bool more = c + 1 < ntiles;
					#pragma unroll
					for (int j = 0; j < 4; j++) cb[j] = pc[T - 12 + j];
					const bool ready = more && kb_mbar_test(&S.b_done[(c + 1) & 3], ((c + 1) >> 2) & 1);
					group(ca, T - 16);
					#pragma unroll
					for (int j = 0; j < 4; j++) ca[j] = pc[T - 8 + j];
					group(cb, T - 12);
					#pragma unroll
					for (int j = 0; j < 4; j++) cb[j] = pc[T - 4 + j];
					group(ca, T - 8);
					if (ready) {
						const float4* pn = S.coef[(c + 1) & 3].r[v];
						#pragma unroll
						for (int j = 0; j < 4; j++) ca[j] = pn[j];
						have_next = true;
					}
					group(cb, T - 4);
				} else {
					for (; t + 8 <= steps; t += 8) {
						#pragma unroll
						for (int j = 0; j < 4; j++) cb[j] = pc[t + 4 + j];
						group(ca, t);
						#pragma unroll
						for (int j = 0; j < 4; j++) ca[j] = pc[t + 8 + j];
						group(cb, t + 4);
					}
					for (; t < steps; t++) {
						const float4 c1 = pc[t];
						const float y = c1.x + z0;
						z0 = c1.y - c1.z * y + z1;
						z1 = c1.x - c1.w * y;
						po[t] = y;
					}
				}
			}
			if (CHECK && is_flt && S.coef[c & 3].r[role_voice][T].x != (float)c) bad = true;
			if (CHECK && bad) S.protocol_error = 1;
			__syncwarp();
			if (lane == 0) { kb_mbar_arrive(&S.c_done[c & 3]); KB_C2_TR(1, c, 1); }
		}
	} else if (worker) {                                                 // ---- D (tile j - LAG) then B (tile j), two rounds each, warp by warp
		const bool pairs_ok = (n & 1) == 0 && (reinterpret_cast<size_t>(dst) & 7) == 0;
		for (int j = 0; j < ntiles + LAG; j++) {
			if (wtid == 0) KB_C2_TR(9, j, 0);
			const int d = j - LAG;
			if (d >= 0) {
				const int steps = min(T, n - d * T);
				kb_mbar_wait(&S.c_done[d & 3], (d >> 2) & 1);
				if (wtid == 0) KB_C2_TR(9, j, 1);
				if (!CHECK && steps == T && pairs_ok) {                          // a full tile: two samples per thread, 64-bit loads and stores
					for (int item = wtid; item < G * (T / 2); item += wthreads) {
						const int v = item / (T / 2), t = (item % (T / 2)) * 2;
						if (v0 + v < total) {
							float2 y = make_float2(0.f, 0.f);
							if (S.c.active[v]) {
								const float2 o = *reinterpret_cast<const float2*>(&S.out[d & 3].r[v][t]), a = *reinterpret_cast<const float2*>(&S.amp[d & 7].r[v][t]);
								y = make_float2(o.x * a.x, o.y * a.y);                   // out *= adsr++   Filter.k:33
							}
							*reinterpret_cast<float2*>(dst + (size_t)(v0 + v) * n + d * T + t) = y;
						}
					}
				} else
				#pragma unroll
				for (int r = 0; r < 2; r++) {
					const int item = wtid + r * wthreads, v = item / T, t = item % T;
					if (t < steps && v0 + v < total) {                           // out *= adsr++   Filter.k:33
						bool bad = false;
						if (CHECK && S.c.active[v]) bad = S.out[d & 3].r[v][T] != (float)d || S.amp[d & 7].r[v][T] != (float)d;
						float y = S.c.active[v] ? S.out[d & 3].r[v][t] * S.amp[d & 7].r[v][t] : 0.f;
						if (CHECK && S.c.active[v]) { __threadfence_block(); bad = bad || S.out[d & 3].r[v][T] != (float)d || S.amp[d & 7].r[v][T] != (float)d || S.protocol_error; }
						if (CHECK && bad) { y = __int_as_float(0x7fc00000); S.protocol_error = 1; }
						dst[(size_t)(v0 + v) * n + d * T + t] = y;
					}
				}
			}
			if (wtid == 0) KB_C2_TR(3, j, 1);
			if (j < ntiles) {
				const int b = j, steps = min(T, n - b * T);
				kb_mbar_wait(&S.a_done[b & 3], (b >> 2) & 1);
				if (wtid == 0) KB_C2_TR(2, j, 0);
				#pragma unroll
				for (int r = 0; r < 2; r++) {
					const int item = wtid + r * wthreads, v = item / T, t = item % T;
					if (t < steps && S.c.active[v]) {
						if (CHECK) { if (S.cut[b & 3].r[v][T] != (float)b) S.protocol_error = 1; if ((t & 31) == 0) { S.coef[b & 3].r[v][T].x = (float)b; } __threadfence_block(); }
						const float f = S.cut[b & 3].r[v][t];
						const float w = f * fs.w;
						float sin0, cos0;
						kb_sincosf(w, sin0, cos0);
						const float a = sin0 / (2.f * 10.f);
						const float inv = kb_const_inv(1.f + a);
						float4 cf;
						cf.z = inv * (-2.f * cos0);
						cf.w = inv * (1.f - a);
						cf.x = inv * (1.f - cos0) * 0.5f;
						cf.y = inv * (1.f - cos0);
						if (b * T + t == n - 1) S.lastc[v] = cf;
						const float in = kb_osm_at(S.osc[v], (uint32_t)(b * T + t));
						cf.x = cf.x * in; cf.y = cf.y * in;
						S.coef[b & 3].r[v][t] = cf;
						if (CHECK && S.cut[b & 3].r[v][T] != (float)b) S.protocol_error = 1;
					}
				}
				__syncwarp();
				if (lane == 0) kb_mbar_arrive(&S.b_done[b & 3]);
				if (wtid == 0) KB_C2_TR(2, j, 1);
			}
		}
	}
	#undef KB_C2_TR
	__syncthreads();
	if (trace && tid == 0 && blockIdx.x < 256) { const long long t1 = clock64(); trace[(4 * 64 + blockIdx.x) * 2] = t_loop - t_entry; trace[(4 * 64 + blockIdx.x) * 2 + 1] = t1 - t_loop; }

	// write the state back
	if (is_env) { kb_envr_store(env, voices[v0 + role_voice].env); voices[v0 + role_voice].filter.f = env.out; voices[v0 + role_voice].filter.Q = 10.f; }
	if (is_adsr) {
		kb_envr_store(env, voices[v0 + role_voice].adsr);
		if (env.stage == KB_ENV_OFF) hdr[v0 + role_voice].stage = KB_NOTE_OFF;
	}
	if (is_flt) {
		KbBiquad& b = voices[v0 + role_voice].filter;
		const float4 lc = S.lastc[role_voice];
		b.z0 = z0; b.z1 = z1; b.b0 = lc.x; b.b2 = lc.x; b.b1 = lc.y; b.a1 = lc.z; b.a2 = lc.w;
	}
	if (worker && wtid < G && S.c.active[wtid]) {
		KbOsm o = S.osc[wtid];
		kb_osm_advance(o, (uint32_t)n);
		voices[v0 + wtid].osc.offset = o.offset;
		voices[v0 + wtid].osc.state = o.state;
	}
}

// ------------------------------------------------------------------------------------------------ SuperSaw
// SuperSaw.k:25-33: seven detuned saws summed (each `/ 7`, in index order) times the ADSR.
//   A (tile k)    warp 0, lane = voice: the ADSR
//   B (tile k-1)  thread = (oscillator, voice, t): one closed-form oscillator sample `/ 7` into shared memory — seven times
//                 the parallelism of a thread per (voice, t), which matters because 256 voices x 128 samples fill only a
//                 quarter of the chip's lanes
//   D (tile k-2)  thread = (voice, t): the seven parts added in oscillator order (the reference's `out +=` order), `* adsr`,
//                 coalesced store
template <int G> struct KbSsawSmem {
	KbTileCommon<G> c;
	KbTileRows<G> amp[4];            // A -> D, two ticks later
	KbTileRows<G> part[2][7];        // B -> D: osc[j] / 7 per (voice, t)
	KbOsm osc[G][7];
};
template <int G, int NT>
__global__ void __launch_bounds__(NT, 1) kb_ssaw_tiled_kernel(KbSsawVoice* __restrict__ voices, KbVoiceHdr* __restrict__ hdr,
                                                                         float* __restrict__ dst, int n, int total, KbFs fs) {
	constexpr int T = KB_TILE_T;
	extern __shared__ __align__(16) unsigned char kb_smem[];
	KbSsawSmem<G>& S = *reinterpret_cast<KbSsawSmem<G>*>(kb_smem);
	const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
	const int v0 = blockIdx.x * G;
	kb_tile_prologue(S.c, voices, hdr, v0, total);
	const bool is_env = warp == 0 && lane < G && S.c.active[lane];
	KbEnvR env;
	if (is_env) kb_tile_load_env(S.c, lane, voices[v0 + lane].adsr, env);
	if (tid >= 32 && tid < 32 + G * 7 && S.c.active[(tid - 32) / 7]) S.osc[(tid - 32) / 7][(tid - 32) % 7] = voices[v0 + (tid - 32) / 7].osc[(tid - 32) % 7];
	__syncthreads();
	const int ntiles = (n + T - 1) / T;
	const int wtid = tid - 32, wthreads = NT - 32;
	for (int k = 0; k < ntiles + 2; k++) {
		if (warp == 0) {                                                     // ---- A, tile k
			if (is_env && k < ntiles) {
				const int steps = min(T, n - k * T);
				kb_envr_run_tile<false>(fs, env, S.c.px[lane], S.c.py[lane], S.amp[k & 3].r[lane], steps);
			}
		} else {
			const int b = k - 1, d = k - 2;
			if (b >= 0 && b < ntiles) {                                      // ---- B, tile k-1
				const int steps = min(T, n - b * T);
				for (int item = wtid; item < 7 * G * T; item += wthreads) {
					const int j = item / (G * T), v = (item / T) % G, t = item % T;
					if (t < steps && S.c.active[v]) S.part[b & 1][j].r[v][t] = kb_osm_at(S.osc[v][j], (uint32_t)(b * T + t)) / 7;
				}
			}
			if (d >= 0 && d < ntiles) {                                      // ---- D, tile k-2
				const int steps = min(T, n - d * T);
				for (int item = wtid; item < G * T; item += wthreads) {
					const int v = item / T, t = item % T;
					if (t < steps && v0 + v < total) {
						float out = 0.f;
						if (S.c.active[v]) {
							#pragma unroll
							for (int j = 0; j < 7; j++) out += S.part[d & 1][j].r[v][t];   // out += osc[j] / 7, j = 0..6  SuperSaw.k:29-31
							out *= S.amp[d & 3].r[v][t];
						}
						dst[(size_t)(v0 + v) * n + d * T + t] = out;
					}
				}
			}
		}
		__syncthreads();
	}
	if (is_env) {
		kb_envr_store(env, voices[v0 + lane].adsr);
		if (env.stage == KB_ENV_OFF) hdr[v0 + lane].stage = KB_NOTE_OFF;
	}
	if (tid >= 32 && tid < 32 + G * 7 && S.c.active[(tid - 32) / 7]) {
		KbOsm o = S.osc[(tid - 32) / 7][(tid - 32) % 7];
		kb_osm_advance(o, (uint32_t)n);
		voices[v0 + (tid - 32) / 7].osc[(tid - 32) % 7].offset = o.offset;
		voices[v0 + (tid - 32) / 7].osc[(tid - 32) % 7].state = o.state;
	}
}

// ------------------------------------------------------------------------------------------------------ FM
// FM.k:61-73 (three Operator<Sine> in series, klang.h:4140-4173).  Nothing in the voice but its four envelopes is a recurrence:
// the operator phases are integer ramps and each operator is phase-modulated by the previous operator's output of the SAME
// sample (kb_fm_at, kb_graphs.cuh).
//   A (tile k)    warp 0, lane = (envelope, voice): the three operator envelopes and the ADSR, run-length form
//   D (tile k-1)  thread = (voice, t): the whole operator chain of one sample, coalesced store
// The block-start voice state sits read-only in shared memory; kb_fm_block_end writes the oscillators back.
// tests/host/fm_tiled_check.cpp proves this formulation (same functions, g++) bit-identical to the per-tick kb_fm_tick,
// samples and state; tools/fm_tiled_probe.py compares the kernel with the lane-per-voice kernel on the device
// (1024 voices x 4096 samples: 58 us against 2502 us).
template <int G> struct KbFmSmem {
	KbFmVoice voice[G];
	KbTileRows<G> env[2][4];         // A -> D: [tile parity][op0, op1, op2, adsr]
	KbFmSample last[G];              // the block's final sample (offsets and operator outputs the next block starts from)
	float i1[G], i2[G];
	int active[G];
};
template <int G, int NT>
__global__ void __launch_bounds__(NT, 1) kb_fm_tiled_kernel(KbFmVoice* __restrict__ voices, KbVoiceHdr* __restrict__ hdr,
                                                                       const KbSynthBlock* __restrict__ blk, float* __restrict__ dst,
                                                                       int n, int voices_per_inst, int total, KbFs fs) {
	static_assert(4 * G <= 32, "the four envelopes of every voice share warp 0");
	constexpr int T = KB_TILE_T;
	extern __shared__ __align__(16) unsigned char kb_smem[];
	KbFmSmem<G>& S = *reinterpret_cast<KbFmSmem<G>*>(kb_smem);
	const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
	const int v0 = blockIdx.x * G;
	if (tid < G) {
		const int v = v0 + tid;
		const int act = (v < total && hdr[v].stage != KB_NOTE_OFF) ? 1 : 0;
		S.active[tid] = act;
		if (v < total) hdr[v].active = act;
		if (act) {
			S.voice[tid] = voices[v];
			S.i1[tid] = blk[v / voices_per_inst].fm_i1;
			S.i2[tid] = blk[v / voices_per_inst].fm_i2;
		}
	}
	__syncthreads();
	const int ev = lane % G, ee = lane / G;                               // A lanes: envelope ee (3 = ADSR) of voice ev
	const bool is_env = warp == 0 && lane < 4 * G && S.active[ev];
	const KbEnv* esrc = nullptr;
	KbEnvR env;
	if (is_env) {
		esrc = ee < 3 ? &S.voice[ev].op[ee].env : &S.voice[ev].adsr;      // breakpoints stay in shared memory
		kb_envr_load(env, *esrc);
	}
	const int ntiles = (n + T - 1) / T;
	const int wtid = tid - 32, wthreads = NT - 32;
	for (int k = 0; k < ntiles + 1; k++) {
		if (warp == 0) {                                                     // ---- A, tile k
			if (is_env && k < ntiles) {
				const int steps = min(T, n - k * T);
				kb_envr_run_tile<false>(fs, env, esrc->px, esrc->py, S.env[k & 1][ee].r[ev], steps);
			}
		} else {
			const int d = k - 1;
			if (d >= 0) {                                                    // ---- D, tile k-1
				const int steps = min(T, n - d * T);
				for (int item = wtid; item < G * T; item += wthreads) {
					const int v = item / T, t = item % T;
					if (t < steps && v0 + v < total) {
						float out = 0.f;
						if (S.active[v]) {
							const KbFmSample s = kb_fm_at(S.voice[v], (uint32_t)(d * T + t), S.i1[v], S.i2[v], S.env[d & 1][0].r[v][t],
							                              S.env[d & 1][1].r[v][t], S.env[d & 1][2].r[v][t], S.env[d & 1][3].r[v][t]);
							out = s.out;
							if (d * T + t == n - 1) S.last[v] = s;
						}
						dst[(size_t)(v0 + v) * n + d * T + t] = out;
					}
				}
			}
		}
		__syncthreads();
	}
	if (is_env) {
		KbEnv& e = ee < 3 ? voices[v0 + ev].op[ee].env : voices[v0 + ev].adsr;
		kb_envr_store(env, e);
		if (ee == 3 && env.stage == KB_ENV_OFF) hdr[v0 + ev].stage = KB_NOTE_OFF;    // FM.k:71-72 `if (adsr.finished()) stop()`
	}
	if (warp == 1 && lane < G && S.active[lane]) {
		KbFmVoice m = S.voice[lane];
		kb_fm_block_end(m, (uint32_t)n, S.i1[lane], S.i2[lane], S.last[lane]);
		KbFmVoice& o = voices[v0 + lane];
		#pragma unroll
		for (int j = 0; j < 3; j++) { o.op[j].osc.position = m.op[j].osc.position; o.op[j].osc.offset = m.op[j].osc.offset; o.op[j].amp = m.op[j].amp; o.op[j].in = m.op[j].in; }
	}
}

// ------------------------------------------------------------------ one envelope x closed-form oscillators
// Breakpoint.k / Ramp.k / Release.k (KbSenvVoice) and Modulation/AM.k (KbSmodVoice): the FM.k kernel with one envelope per voice.
//   A (tile k)    warp 0, lane = voice: the envelope, run-length form
//   D (tile k-1)  thread = (voice, t): kb_es_at, coalesced store
// tests/host/esine_check.cpp proves the formulation (same functions, g++) bit-identical to the per-tick forms, samples and state.
template <class VOICE, int G> struct KbEsSmem {
	VOICE voice[G];                  // block-start state after kb_es_begin (read-only during the block)
	KbTileRows<G> env[2];            // A -> D
	float c1[G];
	int active[G];
};
template <class VOICE, int G, int NT>
__global__ void __launch_bounds__(NT, 1) kb_esine_tiled_kernel(VOICE* __restrict__ voices, KbVoiceHdr* __restrict__ hdr,
                                                                          const KbSynthBlock* __restrict__ blk, float* __restrict__ dst,
                                                                          int n, int voices_per_inst, int total, KbFs fs) {
	static_assert(G <= 32, "one envelope lane per voice in warp 0");
	constexpr int T = KB_TILE_T;
	extern __shared__ __align__(16) unsigned char kb_smem[];
	KbEsSmem<VOICE, G>& S = *reinterpret_cast<KbEsSmem<VOICE, G>*>(kb_smem);
	const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
	const int v0 = blockIdx.x * G;
	if (tid < G) {
		const int v = v0 + tid;
		const int act = (v < total && hdr[v].stage != KB_NOTE_OFF) ? 1 : 0;
		S.active[tid] = act;
		if (v < total) hdr[v].active = act;
		if (act) {
			S.voice[tid] = voices[v];
			S.c1[tid] = blk[v / voices_per_inst].c[1];
			kb_es_begin(fs, S.voice[tid], blk[v / voices_per_inst].c[0]);
		}
	}
	__syncthreads();
	const bool is_env = warp == 0 && lane < G && S.active[lane];
	const KbEnv* esrc = nullptr;
	KbEnvR env;
	if (is_env) {
		esrc = &kb_es_env(S.voice[lane]);                                 // breakpoints stay in shared memory
		kb_envr_load(env, *esrc);
	}
	const int ntiles = (n + T - 1) / T;
	const int wtid = tid - 32, wthreads = NT - 32;
	for (int k = 0; k < ntiles + 1; k++) {
		if (warp == 0) {                                                     // ---- A, tile k
			if (is_env && k < ntiles) {
				const int steps = min(T, n - k * T);
				kb_envr_run_tile<false>(fs, env, esrc->px, esrc->py, S.env[k & 1].r[lane], steps);
			}
		} else {
			const int d = k - 1;
			if (d >= 0) {                                                    // ---- D, tile k-1
				const int steps = min(T, n - d * T);
				for (int item = wtid; item < G * T; item += wthreads) {
					const int v = item / T, t = item % T;
					if (t < steps && v0 + v < total)
						dst[(size_t)(v0 + v) * n + d * T + t] = S.active[v] ? kb_es_at(S.voice[v], (uint32_t)(d * T + t), S.c1[v], S.env[d & 1].r[v][t]) : 0.f;
				}
			}
		}
		__syncthreads();
	}
	if (is_env) {
		kb_envr_store(env, kb_es_env(voices[v0 + lane]));
		if (kb_es_stops(S.voice[lane]) && env.stage == KB_ENV_OFF) hdr[v0 + lane].stage = KB_NOTE_OFF;
	}
	if (warp == 1 && lane < G && S.active[lane]) {
		VOICE m = S.voice[lane];
		kb_es_end(m, (uint32_t)n);
		kb_es_writeback(voices[v0 + lane], m);
	}
}

// --------------------------------------------------------------------------------------------------- TB303
// TB303.k:103-113.  A: filter envelope and ADSR.  B: oscillator sample and Filter::set coefficients b0, k, g
// (TB303.k:37-55; polynomials of the cutoff, no transcendental: r and the feedback one-pole are control-rate
// constants from the host, KbTbBlock).  C: the ladder recurrence with its one-pole feedback (TB303.k:70-79), which
// parks g*z[3] per sample.  D: the soft clip (tanhf, double divide), `* adsr`, stores.
// Filter::set only recomputes when (cutoff, resonance, drive) change; its outputs are pure functions of those three, so
// B evaluates them for every sample and C adopts them exactly when the reference's comparison says "changed".
template <int G> struct KbTbSmem {
	KbTileCommon<G> c;
	KbTileRows<G> e[2], amp[4], b0[2], kk[2], g[2], x[2], cut[2], y[2];
	KbOsm osc[G];
	KbTbBlock blk[G];
	float vf[G], last_out[G];
};
template <int G, int NT, int LAYOUT = 0>
__global__ void __launch_bounds__(NT, 1) kb_tb_tiled_kernel(KbTbVoice* __restrict__ voices, KbVoiceHdr* __restrict__ hdr,
                                                                       const KbSynthBlock* __restrict__ blk, float* __restrict__ dst,
                                                                       int n, int voices_per_inst, int total, KbFs fs) {
	constexpr int T = KB_TILE_T;
	extern __shared__ __align__(16) unsigned char kb_smem[];
	KbTbSmem<G>& S = *reinterpret_cast<KbTbSmem<G>*>(kb_smem);
	const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
	const int v0 = blockIdx.x * G;
	kb_tile_prologue(S.c, voices, hdr, v0, total);
	const int role = kb_tile_role<LAYOUT>(warp);                      // 0 = A (filter envelope), 1 = A (ADSR), 2 = C (ladder), -1 = worker / idle
	constexpr bool MERGED_A = LAYOUT == 2 && 2 * G <= 32;             // role 0 runs both envelopes of a voice on lanes v and G + v
	const bool worker = kb_tile_is_worker<LAYOUT>(warp);
	const bool first_worker = worker && kb_tile_worker_tid<LAYOUT, NT>(warp, lane) < 32;
	if (first_worker && lane < G && v0 + lane < total) {
		S.blk[lane] = blk[(v0 + lane) / voices_per_inst].tb;
		if (S.c.active[lane]) {
			S.osc[lane] = S.blk[lane].is_square ? voices[v0 + lane].square : voices[v0 + lane].saw;
			S.vf[lane] = voices[v0 + lane].f;
			S.last_out[lane] = voices[v0 + lane].filter.out;
		}
	}
	const int a_sub = (MERGED_A && role == 0 && lane >= G) ? 1 : 0;
	const int role_voice = lane - a_sub * G;
	const bool role_ok = role_voice < G && S.c.active[role_voice < G ? role_voice : 0];
	const bool is_env = role == 0 && a_sub == 0 && role_ok, is_adsr = (role == 1 || a_sub == 1) && role_ok, is_flt = role == 2 && role_ok;
	const int slot = (is_adsr ? G : 0) + role_voice;
	KbEnvR env;
	if (is_env) kb_tile_load_env(S.c, slot, voices[v0 + role_voice].env, env);
	if (is_adsr) kb_tile_load_env(S.c, slot, voices[v0 + role_voice].adsr, env);
	KbTbFilter F;
	if (is_flt) F = voices[v0 + role_voice].filter;
	__syncthreads();
	const int ntiles = (n + T - 1) / T;
	const int wtid = kb_tile_worker_tid<LAYOUT, NT>(warp, lane);
	constexpr int wthreads = kb_tile_worker_threads_v<LAYOUT, NT>;
	static_assert(LAYOUT != 2 || NT % 128 == 0, "layout 2 needs whole rows of four warps");
	for (int k = 0; k < ntiles + 3; k++) {
		if (role == 0 || role == 1) {                                    // ---- A, tile k
			if ((is_env || is_adsr) && k < ntiles) {
				const int steps = min(T, n - k * T);
				float* row = is_env ? S.e[k & 1].r[role_voice] : S.amp[k & 3].r[role_voice];
				kb_envr_run_tile<false>(fs, env, S.c.px[slot], S.c.py[slot], row, steps);
			}
		} else if (role == 2) {                                          // ---- C, tile k-2
			const int c = k - 2;
			if (is_flt && c >= 0 && c < ntiles) {
				const int steps = min(T, n - c * T), v = role_voice;
				const KbTbBlock B = S.blk[v];
				const float* pcut = S.cut[c & 1].r[v]; const float* pb0 = S.b0[c & 1].r[v]; const float* pkk = S.kk[c & 1].r[v];
				const float* pg = S.g[c & 1].r[v]; const float* px_ = S.x[c & 1].r[v];
				// the operands of the NEXT sample are loaded before this sample's dependent chain (rows are padded by one)
				float n_cut = pcut[0], n_b0 = pb0[0], n_kk = pkk[0], n_g = pg[0], n_x = px_[0];
				for (int t = 0; t < steps; t++) {
					const float cutoff = n_cut, o_b0 = n_b0, o_kk = n_kk, o_g = n_g, o_x = n_x;
					n_cut = pcut[t + 1]; n_b0 = pb0[t + 1]; n_kk = pkk[t + 1]; n_g = pg[t + 1]; n_x = px_[t + 1];
					if (F.cutoff != cutoff || F.resonance != B.resonance || F.drive != B.drive) {
						F.cutoff = cutoff; F.resonance = B.resonance; F.drive = B.drive; F.r = B.r;
						F.b0 = o_b0; F.k = o_kk; F.g = o_g;
					}
					if (F.feedback.f != B.hpf_f) { F.feedback.f = B.hpf_f; F.feedback.b0 = B.hpf_b0; F.feedback.b1 = B.hpf_b1; F.feedback.a1 = B.hpf_a1; }   // setHPF  TB303.k:33-35
					F.in = o_x;
					const float y0 = kb_onepole_tick(F.feedback, F.k * F.z[3]) * 0.9f * F.resonance;
					const float shaped = (y0 > KB_ROOT2_F) ? KB_ROOT2_F : (y0 < -0.5f) ? -0.5f : y0;   // (float)y0 < -0.5 (double literal) == y0 < -0.5f: -0.5 is exact
					F.in -= shaped;
					F.z[0] += 2.f * F.b0 * (F.in - F.z[0] + F.z[1]);
					F.z[1] += F.b0 * (F.z[0] - 2.f * F.z[1] + F.z[2]);
					F.z[2] += F.b0 * (F.z[1] - 2.f * F.z[2] + F.z[3]);
					F.z[3] += F.b0 * (F.z[2] - 2.f * F.z[3]);
					S.y[c & 1].r[v][t] = F.g * F.z[3];
				}
			}
		} else if (worker) {
			const int b = k - 1, d = k - 3;
			if (b >= 0 && b < ntiles) {                                      // ---- B, tile k-1
				const int steps = min(T, n - b * T);
				for (int item = wtid; item < G * T; item += wthreads) {
					const int v = item / T, t = item % T;
					if (t < steps && S.c.active[v]) {
						const KbTbBlock& B = S.blk[v];
						const float e = S.e[b & 1].r[v][t];
						float cutoff = (S.vf[v] + B.c0sq_nyq) * (e * e);                 // TB303.k:106
						if (cutoff > fs.nyquist) cutoff = fs.nyquist;
						const float fx = cutoff * fs.inv * KB_ROOT2_INV_F;
						S.b0[b & 1].r[v][t] = (0.00045522346f + 6.1922189f * fx) / (1.f + 12.358354f * fx + 4.4156345f * (fx * fx));
						float kq = fx*(fx*(fx*(fx*(fx*(fx+7198.6997f)-5837.7917f)-476.47308f)+614.95611f)+213.87126f)+16.998792f;
						float g = kq * 0.058823529411764705882352941176471f;
						g = (g - 1.f) * B.r + 1.f;
						g = (g * (1.f + B.r));
						kq = kq * B.r;
						S.kk[b & 1].r[v][t] = kq; S.g[b & 1].r[v][t] = g;
						S.cut[b & 1].r[v][t] = cutoff;
						const float osc = kb_osm_at(S.osc[v], (uint32_t)(b * T + t));
						S.x[b & 1].r[v][t] = B.is_square ? osc * 0.5f : osc;
					}
				}
			}
			if (d >= 0 && d < ntiles) {                                      // ---- D, tile k-3
				const int steps = min(T, n - d * T);
				for (int item = wtid; item < G * T; item += wthreads) {
					const int v = item / T, t = item % T;
					if (t < steps && v0 + v < total) {
						float out = 0.f;
						if (S.c.active[v]) {
							const float x = S.y[d & 1].r[v][t];
							const float hard = (x > 1.f) ? 1.f : (x < -1.f) ? -1.f : x;
							const float clipped = (float)((double)kb_tanhf(hard * S.blk[v].drive) / S.blk[v].clip_den);   // TB303.k:66-68
							if (d * T + t == n - 1) S.last_out[v] = clipped;                                              // Filter::out after the block
							out = clipped * S.amp[d & 3].r[v][t];
						}
						dst[(size_t)(v0 + v) * n + d * T + t] = out;
					}
				}
			}
		}
		__syncthreads();
	}
	if (is_env) kb_envr_store(env, voices[v0 + role_voice].env);
	if (is_adsr) {
		kb_envr_store(env, voices[v0 + role_voice].adsr);
		if (env.stage == KB_ENV_OFF) hdr[v0 + role_voice].stage = KB_NOTE_OFF;
	}
	if (is_flt) { F.out = S.last_out[role_voice]; voices[v0 + role_voice].filter = F; }
	if (first_worker && lane < G && S.c.active[lane]) {
		KbOsm o = S.osc[lane];
		kb_osm_advance(o, (uint32_t)n);
		KbOsm& dsto = S.blk[lane].is_square ? voices[v0 + lane].square : voices[v0 + lane].saw;
		dsto.offset = o.offset; dsto.state = o.state;
	}
}

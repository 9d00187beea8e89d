// klang-b200 — Reverb.k (examples/Reverb.k:9-279), third schedule: decoupled warp roles, delay lines RESIDENT in shared memory.
//
// Same arithmetic as kb_reverb_par_kernel / kb_reverb_pipe_kernel (kb_fx_parallel.cuh) and as the frame-sequential kb_reverb_frame: one
// CTA per (instance, side), time cut into chunks of at most a QUARTER of the shortest read-to-write distance of the 8 feedback lines, so
// the ring window of chunk k is complete once chunk k-2 has been written.  What changes is where the data lives and how the roles meet:
//
//   * the LIVE SPAN of every delay line — the samples between its read head and its write head, 7 .. 34 ms each, ~30 KB for the 8 lines
//     of a side at 48 kHz — and the whole early-reflection ring of the side (Stereo::Delay<21600>, 86.4 KB) are resident in shared memory
//     for the launch: they arrive once by 1-D cp.async.bulk (SASS UBLKCP) completing on an mbarrier (SYNCS), as do the io chunks.  In
//     the steady state a chunk touches HBM only for its io (one bulk copy in, coalesced stores out) and for the write-through of the
//     ring samples it produces (stores nobody in this launch reads back): no load sits between two chunks of the feedback loop, and the
//     LSU the filter warp shares with the rest of its SM carries stores only.  (Round 1 gathered every window and every tap from L2.)
//   * no CTA-wide barrier per chunk.  Every role runs its own loop over the chunks and hands over through progress counters in shared
//     memory (release / acquire fences around a volatile word).  The slow role never waits for the latency of the fast ones.
//   * the filter lane receives PRE-MULTIPLIED operands (b0 x, b1 x, b2 x) from worker warps, so its own instruction stream is the bare
//     recurrence  y = p0 + z0;  z0 = (p1 - a1 y) + z1;  z1 = p2 - a2 y  — 6 operations per tick, 4 of them on the dependent chain
//     (the same roundings as Biquad::Filter::process, klang.h:5605-5612: every product and sum is rounded on its own).
//   An instance whose live spans do not fit (fs = 192 kHz) keeps the round-1 pipeline (kb_reverb_pipe_kernel), chosen per instance by the plan.
//
// KB_FX_TOLERANCE (opt-in, include/klang_b200.h) replaces the serial filter lane by a warp-per-line parallel scan (kb_rv3_scan_chunk): each
// lane runs 5 consecutive ticks, the lane-end states are combined by a Kogge-Stone scan over powers of the 2x2 state-transition matrix,
// each lane re-runs its ticks from its true start state.  That RE-ASSOCIATES the fp32 recurrence, so results are no longer bit-identical:
// the plan admits an instance only when the rounding-noise gain of its line filters keeps the output within the parity bar
// 1e-5 |r| + 1e-6 peak of the reference (profiles/r02_reverb_tolerance.txt; the early LPF -> HPF cascade, whose 100 Hz high-pass
// amplifies rounding noise 40x above that bar, stays serial and exact in both modes).
#pragma once
#include "kb_fx_parallel.cuh"
#include "kb_scan.cuh"
#include "kb_sync.cuh"

#define KB_RV3_LMAX 80                        // frames per chunk (a multiple of 4: io chunks are 16-byte bulk copies)
#define KB_RV3_XROW 161                       // float4 per operand row (odd: the 8 lanes of the filter warp hit distinct banks)
#define KB_RV3_YROW 164                       // floats per output row (164 = 4 mod 32: conflict-free 128-bit stores from 8 lanes)
#define KB_RV3_EROW (KB_RV3_LMAX + 16)        // early rows: LMAX frames + the read-ahead of the row filter
#define KB_RV3_DI 8                           // io chunks in flight
#define KB_RV3_DE 4                           // early-cascade output chunks in flight
#define KB_RV3_DR 4                           // early-reflection rows (r1) in flight
#define KB_RV3_ESIZE 21600                    // Stereo::Delay<21600>  Reverb.k:11
#define KB_RV3_LRCAP 20480                    // floats of shared memory for the live spans of the 8 lines of a side
#define KB_RV3_NT 640                         // threads, exact mode (20 warps)
#define KB_RV3_NT_TOL 768                     // threads, tolerance mode (24 warps)

struct KbRv3LinePlan { int lag, cap, scan_ok, rpos; };
struct KbRv3Smem {
	float er[KB_RV3_ESIZE];                   // the early ring of this side, resident for the launch
	float lr[KB_RV3_LRCAP];                   // the live spans of the 8 feedback lines: line l = a ring of lcap[l] floats at lbase[l]
	float4 xq[2][8][KB_RV3_XROW];             // exact mode: filter operands per tick (b0 x, b1 x, b2 x, -), double buffered
	float y[2][8][KB_RV3_YROW];               // filter outputs per tick, double buffered
	float xin[KB_RV3_DI][KB_RV3_EROW];        // io chunks (bulk copies)
	float ylp[2][KB_RV3_EROW];                // early LPF output
	float xf[KB_RV3_DE][KB_RV3_EROW];         // early LPF -> HPF output
	float r1[KB_RV3_DR][KB_RV3_LMAX];         // early reflections per frame (taps summed in order)
	float carry[2][8];                        // FilteredDelay::in carried between frames and chunks: [old/new][line]
	float times[KB_RV_MAXREFL], gg[KB_RV_MAXREFL];
	long long lring[8]; int lsize[8], rpos0[8], wpos0[8];
	int lbase[8], lcap[8], lro[8], lwo[8];    // shared-memory ring of line l: base, capacity, offsets of the read / write head at block start
	float frac[8], gain[8], b0[8], b1[8], b2[8], a1[8], a2[8];
	unsigned long long bar_xin[KB_RV3_DI], bar_res;                 // mbarriers: bulk-copy completion (io chunks; the resident rings)
	int p_done, f_done, w_done, t_done, e_done;                     // chunks completed per role
	int f_cnt[2];                                                   // tolerance mode: line-chunks completed, per chunk parity (8 per chunk)
	KbRv3LinePlan pline[16]; KbFxPlan plan;                         // the plan of this instance, made by the CTA itself
};

// 1-D bulk copy global -> shared, completing `bytes` on the mbarrier.  Addresses and size are multiples of 16 bytes.
KB_D void kb_bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
	             :: "r"(kb_smem_u32(dst)), "l"(src), "r"(bytes), "r"(kb_smem_u32(bar)) : "memory");
}
KB_D void kb_fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }

// ---- plan: chunk length (a multiple of 4 frames) and the schedule.  mode bits: KB_PLAN_PARALLEL; KB_PLAN_SCAN_OK = the scan is admissible
// for every line; KB_PLAN_RESIDENT = the live spans of both sides fit in shared memory (this file's kernel; otherwise kb_reverb_pipe_kernel)
enum { KB_PLAN_SCAN_OK = 2, KB_PLAN_RESIDENT = 4 };
// shared-memory ring of one line: holds the global ring indices a0 = (read head & ~3) .. write head + one chunk, i.e. lag + 2 LMAX + slack floats
KB_HD int kb_rv3_line_cap(int lag) { return (lag + 2 * KB_RV3_LMAX + 8 + 3) & ~3; }
// one line's contribution (any thread), and the combination (one thread)
KB_HD KbRv3LinePlan kb_rv3_plan_line(const KbRvFDelay& fd) {
	const KbDelay& d = fd.delay;
	KbRv3LinePlan r;
	r.lag = d.position - d.last_position; if (r.lag <= 0) r.lag += d.SIZE;      // write head minus read head, in ring samples
	r.cap = kb_rv3_line_cap(r.lag);
	r.scan_ok = kb_rv3_scan_admissible(fd.filter) ? 1 : 0;
	r.rpos = d.last_position;
	return r;
}
KB_HD KbFxPlan kb_rv3_plan_combine(const KbRv3LinePlan* lines, const float* times, int count, int esize_l, int esize_r) {
	int chunk = KB_RV3_LMAX, need[2] = { 0, 0 };
	bool scan_ok = true;
	for (int line = 0; line < 16; line++) {
		chunk = chunk < (lines[line].lag - 2) / 4 ? chunk : (lines[line].lag - 2) / 4;   // two ticks per frame, window of chunk k closed by chunk k-2
		scan_ok = scan_ok && lines[line].scan_ok;
		need[(line >> 2) & 1] += lines[line].cap;                               // lines 0-3, 8-11 belong to side 0 (mid[0], late[0]); 4-7, 12-15 to side 1
	}
	float tmin = 1e30f;
	for (int r = 0; r < count; r++) tmin = fminf(tmin, times[r]);
	chunk = chunk < (int)tmin - 3 ? chunk : (int)tmin - 3;
	chunk &= ~3;
	KbFxPlan p;
	p.chunk = chunk;
	const bool resident = need[0] <= KB_RV3_LRCAP && need[1] <= KB_RV3_LRCAP && esize_l == KB_RV3_ESIZE && esize_r == KB_RV3_ESIZE;
	p.mode = chunk >= 8 ? (KB_PLAN_PARALLEL | (resident ? KB_PLAN_RESIDENT | (scan_ok ? KB_PLAN_SCAN_OK : 0) : 0)) : KB_PLAN_SEQUENTIAL;
	p.gain = p.delay = p.dry = 0.f;
	return p;
}
// (stand-alone form: used when the kernels below are not launched, and by tests of the plan itself)
__global__ void kb_reverb_plan3_kernel(const KbReverb* __restrict__ states, KbFxPlan* __restrict__ plan, int instances) {
	const int inst = blockIdx.x * blockDim.x + threadIdx.x;
	if (inst >= instances) return;
	KbReverb& rv = const_cast<KbReverb&>(states[inst]);
	KbRv3LinePlan lines[16];
	for (int line = 0; line < 16; line++) lines[line] = kb_rv3_plan_line(kb_rv_line(rv, line));
	plan[inst] = kb_rv3_plan_combine(lines, rv.times, rv.count, rv.dl.SIZE, rv.dr.SIZE);
}

// Biquad::Filter::process (klang.h:5605-5612) over pre-multiplied operands, strictly in order, by ONE thread: per tick
//   y = b0 x + z0;  z0' = b1 x - a1 y + z1;  z1' = b2 x - a2 y        with q = (b0 x, b1 x, b2 x)
// Two register sets of 8 ticks used alternately (no copies): while one set is consumed — 8 x 4 dependent operations — the other is loaded, so no
// shared-memory latency sits on the recurrence.  Rows are padded: reads up to 16 ticks past `ticks` stay inside the row's allocation and are unused.
#define KB_RV3_TICK(P, O) { O = P.x + z0; z0 = (P.y - a1 * O) + z1; z1 = P.z - a2 * O; }
KB_D void kb_rv3_filter_row(const float4* __restrict__ q, float* __restrict__ yr, int ticks, float a1, float a2, float& z0, float& z1) {
	float4* y4 = reinterpret_cast<float4*>(yr);
	float4 A0 = q[0], A1 = q[1], A2 = q[2], A3 = q[3], A4 = q[4], A5 = q[5], A6 = q[6], A7 = q[7];
	float4 B0, B1, B2, B3, B4, B5, B6, B7;
	int f = 0;
	for (; f + 16 <= ticks; f += 16) {
		float4 o;
		B0 = q[f + 8]; B1 = q[f + 9]; B2 = q[f + 10]; B3 = q[f + 11]; B4 = q[f + 12]; B5 = q[f + 13]; B6 = q[f + 14]; B7 = q[f + 15];
		KB_RV3_TICK(A0, o.x) KB_RV3_TICK(A1, o.y) KB_RV3_TICK(A2, o.z) KB_RV3_TICK(A3, o.w) y4[(f >> 2)] = o;
		KB_RV3_TICK(A4, o.x) KB_RV3_TICK(A5, o.y) KB_RV3_TICK(A6, o.z) KB_RV3_TICK(A7, o.w) y4[(f >> 2) + 1] = o;
		A0 = q[f + 16]; A1 = q[f + 17]; A2 = q[f + 18]; A3 = q[f + 19]; A4 = q[f + 20]; A5 = q[f + 21]; A6 = q[f + 22]; A7 = q[f + 23];
		KB_RV3_TICK(B0, o.x) KB_RV3_TICK(B1, o.y) KB_RV3_TICK(B2, o.z) KB_RV3_TICK(B3, o.w) y4[(f >> 2) + 2] = o;
		KB_RV3_TICK(B4, o.x) KB_RV3_TICK(B5, o.y) KB_RV3_TICK(B6, o.z) KB_RV3_TICK(B7, o.w) y4[(f >> 2) + 3] = o;
	}
	if (f + 8 <= ticks) {                                            // (set A holds ticks f .. f+7)
		float4 o;
		KB_RV3_TICK(A0, o.x) KB_RV3_TICK(A1, o.y) KB_RV3_TICK(A2, o.z) KB_RV3_TICK(A3, o.w) y4[(f >> 2)] = o;
		KB_RV3_TICK(A4, o.x) KB_RV3_TICK(A5, o.y) KB_RV3_TICK(A6, o.z) KB_RV3_TICK(A7, o.w) y4[(f >> 2) + 1] = o;
		f += 8;
	}
	for (; f < ticks; f++) {
		const float4 p = q[f];
		float y;
		KB_RV3_TICK(p, y)
		yr[f] = y;
	}
}

#ifdef __CUDACC__
// Tolerance mode: one line, one chunk, by one full warp (kb_scan.cuh).  ring / cap / rb = the line's shared-memory ring, its capacity and the
// index of the read head's sample; frac = Delay::process interpolation weight (klang.h:3461-3473); yr = output row.  (z0, z1) = the line's
// filter state, identical in all lanes on entry and on exit.
KB_D void kb_rv3_scan_chunk(const KbRv3ScanCoef& c, const float* __restrict__ ring, int cap, int rb, float frac, int ticks, float* __restrict__ yr, float& z0, float& z1, int lane) {
	const int t0 = lane * KB_RV3_SCAN_P;
	const int cnt = max(0, min(KB_RV3_SCAN_P, ticks - t0));
	float x[KB_RV3_SCAN_P];
	{
		int i = rb + t0; if (i >= cap) i -= cap;
		float wa = cnt > 0 ? ring[i] : 0.f;
		#pragma unroll
		for (int j = 0; j < KB_RV3_SCAN_P; j++) {
			if (++i >= cap) i -= cap;
			const float wb = j < cnt ? ring[i] : 0.f;
			x[j] = wa + frac * (wb - wa);
			wa = wb;
		}
	}
	float e0 = lane == 0 ? z0 : 0.f, e1 = lane == 0 ? z1 : 0.f;
	kb_rv3_scan_run(c, x, cnt, e0, e1, nullptr);
	#pragma unroll
	for (int j = 0; j < 5; j++) {
		const float u0 = __shfl_up_sync(0xffffffffu, e0, 1 << j), u1 = __shfl_up_sync(0xffffffffu, e1, 1 << j);
		if (lane >= (1 << j)) {
			e0 = kb_fma(c.T[j][0], u0, kb_fma(c.T[j][1], u1, e0));
			e1 = kb_fma(c.T[j][2], u0, kb_fma(c.T[j][3], u1, e1));
		}
	}
	// true start state of this lane = true end state of the lane before it (lanes past the last tick are never used)
	float s0 = __shfl_up_sync(0xffffffffu, e0, 1), s1 = __shfl_up_sync(0xffffffffu, e1, 1);
	if (lane == 0) { s0 = z0; s1 = z1; }
	float yv[KB_RV3_SCAN_P];
	kb_rv3_scan_run(c, x, cnt, s0, s1, yv);
	#pragma unroll
	for (int j = 0; j < KB_RV3_SCAN_P; j++) if (j < cnt) yr[t0 + j] = yv[j];
	const int last = (ticks - 1) / KB_RV3_SCAN_P;                    // the lane that holds the chunk's last tick ends in the chunk's end state
	z0 = __shfl_sync(0xffffffffu, s0, last); z1 = __shfl_sync(0xffffffffu, s1, last);
}

// role of a warp.  The warp scheduler of an SM sub-partition (warp id mod 4) favours HIGHER warp ids, so a serial role whose latency bounds the
// kernel is either alone on its sub-partition or holds the highest id there.
// Exact mode (20 warps): the filter warp F (warp 0) is ALONE on sub-partition 0 (warps 4, 8, 12, 16 exit at once; sharing it with the early
//   cascade was measured: F 22 -> 26 cycles per tick); E is the highest warp of sub-partition 1, above one warp each of T, P, W; the rest of the
//   parallel roles are spread evenly over sub-partitions 2 and 3, the roles of the F -> W -> P -> F loop (W, P) above the taps (T).
// Tolerance mode (24 warps): E — now the longest serial chain — is alone on sub-partition 1; the 8 scan warps S hold the highest ids of the others.
enum { KB_RV3_IDLE = 0, KB_RV3_F, KB_RV3_E, KB_RV3_M, KB_RV3_P, KB_RV3_T, KB_RV3_W, KB_RV3_S };
template <int MODE> KB_D void kb_rv3_role(int warp, int& role, int& slot) {
	if (MODE == 0) {
		//                       0            1         2         3            (sub-partition = column)
		const int r[20] = { KB_RV3_F,    KB_RV3_T, KB_RV3_M, KB_RV3_IDLE,
		                    KB_RV3_IDLE, KB_RV3_P, KB_RV3_T, KB_RV3_T,
		                    KB_RV3_IDLE, KB_RV3_W, KB_RV3_P, KB_RV3_P,
		                    KB_RV3_IDLE, KB_RV3_IDLE, KB_RV3_W, KB_RV3_W,
		                    KB_RV3_IDLE, KB_RV3_E, KB_RV3_W, KB_RV3_W };
		const int q[20] = { 0, 0, 0, 0,  0, 0, 1, 2,  0, 0, 1, 2,  0, 0, 1, 2,  0, 0, 3, 4 };       // index of the warp inside its role
		role = r[warp % 20]; slot = q[warp % 20];
	} else {
		const int r[24] = { KB_RV3_T, KB_RV3_IDLE, KB_RV3_M, KB_RV3_T,
		                    KB_RV3_W, KB_RV3_IDLE, KB_RV3_T, KB_RV3_W,
		                    KB_RV3_W, KB_RV3_IDLE, KB_RV3_W, KB_RV3_W,
		                    KB_RV3_S, KB_RV3_IDLE, KB_RV3_S, KB_RV3_S,
		                    KB_RV3_S, KB_RV3_IDLE, KB_RV3_S, KB_RV3_S,
		                    KB_RV3_S, KB_RV3_E,    KB_RV3_S, KB_RV3_IDLE };
		const int q[24] = { 0, 0, 0, 2,  0, 0, 1, 2,  3, 0, 1, 4,  0, 0, 3, 6,  1, 0, 4, 7,  2, 0, 5, 0 };
		role = r[warp % 24]; slot = q[warp % 24];
	}
}

// Measurement aid (KB_RV3_TRACE=<file>, tools/rv3_trace.py): CTA 0 records clock64() when each role starts (after its waits) and ends the work of
// each chunk: trace[((role * KB_RV3_TRACE_CHUNKS + chunk) * 2 + phase)].  nullptr in normal runs.
#define KB_RV3_TRACE_CHUNKS 64
#define KB_RV3_TRACE_ROLES 16
#define KB_RV3_TR(role_, k_, phase_) do { if (trace && blockIdx.x == 0 && (k_) < KB_RV3_TRACE_CHUNKS) trace[(((role_) * KB_RV3_TRACE_CHUNKS + (k_)) * 2 + (phase_))] = clock64(); } while (0)
// skip_scan_ok (exact kernel only): this launch leaves the instances the scan admits to the tolerance kernel launched beside it
template <int MODE>
__global__ void __launch_bounds__(MODE == 0 ? KB_RV3_NT : KB_RV3_NT_TOL) kb_reverb3_kernel(const KbFxHdr* __restrict__ hdrs, KbReverb* __restrict__ states, KbFxPlan* __restrict__ plan,
                                                                                           float* __restrict__ rings, float* __restrict__ io, int n, int stride, int skip_scan_ok, long long* __restrict__ trace) {
	extern __shared__ __align__(128) unsigned char kb_rv3_smem_raw[];
	KbRv3Smem& S = *reinterpret_cast<KbRv3Smem*>(kb_rv3_smem_raw);
	const int inst = blockIdx.x >> 1, side = blockIdx.x & 1;
	const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
	KbReverb& rv = states[inst];
	// the plan of the instance, made here (no plan launch): 16 threads read one line each, thread 0 combines.  It depends only on quantities a
	// block leaves unchanged (read-to-write distances, filter coefficients, tap times), so both CTAs of an instance — and the kernels launched
	// after this one, which read plan[inst] — agree even when one side has already finished.
	if (tid < KB_RV_MAXREFL) { S.times[tid] = rv.times[tid]; S.gg[tid] = side ? rv.gr[tid] : rv.gl[tid]; }
	if (tid >= 32 && tid < 48) S.pline[tid - 32] = kb_rv3_plan_line(kb_rv_line(rv, tid - 32));
	__syncthreads();
	if (tid == 0) {
		S.plan = kb_rv3_plan_combine(S.pline, S.times, rv.count, rv.dl.SIZE, rv.dr.SIZE);
		if (side == 0) plan[inst] = S.plan;
	}
	__syncthreads();
	const KbFxPlan pl = S.plan;
	if (!(pl.mode & KB_PLAN_PARALLEL) || !(pl.mode & KB_PLAN_RESIDENT)) return;
	if (MODE == 1 && !(pl.mode & KB_PLAN_SCAN_OK)) return;          // (the exact kernel launched beside this one takes those)
	if (MODE == 0 && (pl.mode & KB_PLAN_SCAN_OK) && skip_scan_ok) return;
	int role, rslot;
	kb_rv3_role<MODE>(warp, role, rslot);
	const KbControl* c = hdrs[inst].controls;
	float* X = io + ((size_t)inst * 2 + side) * stride;
	const int Lc = pl.chunk, K = (n + Lc - 1) / Lc;
	KbDelay& ed = side ? rv.dr : rv.dl;
	const int epos0 = ed.position;
	float* ringe = rings + ed.ring;
	auto chunk_len = [&](int k) { return min(Lc, n - k * Lc); };

	// ---- prologue: per-line geometry and coefficients, mbarriers, counters
	if (tid < 8) {
		const KbRvFDelay& d = kb_rv_side_line(rv, side, tid);
		S.carry[0][tid] = d.in; S.carry[1][tid] = d.in;
		S.lring[tid] = d.delay.ring; S.lsize[tid] = d.delay.SIZE; S.rpos0[tid] = d.delay.last_position; S.wpos0[tid] = d.delay.position;
		S.frac[tid] = d.delay.last_fraction; S.gain[tid] = d.gain;
		S.b0[tid] = d.filter.b0; S.b1[tid] = d.filter.b1; S.b2[tid] = d.filter.b2; S.a1[tid] = d.filter.a1; S.a2[tid] = d.filter.a2;
	}
	if (tid == 32) {
		// the shared-memory rings: line l holds the global ring indices from a0 = read head & ~3 on; offset o (from a0) lives at lbase + o mod lcap
		int base = 0;
		for (int l = 0; l < 8; l++) {
			const KbRv3LinePlan& lp = S.pline[(l < 4 ? 0 : 8) + side * 4 + (l & 3)];   // line l of this side in kb_rv_line order
			S.lbase[l] = base; S.lcap[l] = lp.cap;
			S.lro[l] = lp.rpos & 3;                                             // read head relative to a0 = read head & ~3
			S.lwo[l] = S.lro[l] + lp.lag;
			base += lp.cap;
		}
		kb_mbar_init(&S.bar_res, 1);
		for (int i = 0; i < KB_RV3_DI; i++) kb_mbar_init(&S.bar_xin[i], 1);
		S.p_done = S.f_done = S.w_done = S.t_done = S.e_done = 0; S.f_cnt[0] = S.f_cnt[1] = 0;
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	__syncthreads();
	if (role == KB_RV3_IDLE) return;

	if (role == KB_RV3_M) {
		// ================================================================ M: the bulk-copy issuer (one thread)
		if (lane != 0) return;
		// resident data, once: the early ring (whole) and the live span of every line, as 16-byte aligned spans (two copies where the global ring wraps)
		unsigned bytes = 4u * KB_RV3_ESIZE;
		for (int l = 0; l < 8; l++) bytes += 4u * (unsigned)((S.lwo[l] + 3) & ~3);
		kb_mbar_expect_tx(&S.bar_res, bytes);
		for (int q = 0; q < 4; q++) kb_bulk_g2s(&S.er[q * (KB_RV3_ESIZE / 4)], ringe + q * (KB_RV3_ESIZE / 4), KB_RV3_ESIZE, &S.bar_res);
		for (int l = 0; l < 8; l++) {
			const float* ring = rings + S.lring[l];
			const int size = S.lsize[l], a0 = S.rpos0[l] & ~3, n4 = (S.lwo[l] + 3) & ~3;
			float* dst = &S.lr[S.lbase[l]];
			if (a0 + n4 <= size) kb_bulk_g2s(dst, ring + a0, 4u * n4, &S.bar_res);
			else {
				const int first = size - a0;
				kb_bulk_g2s(dst, ring + a0, 4u * first, &S.bar_res);
				kb_bulk_g2s(dst + first, ring, 4u * (n4 - first), &S.bar_res);
			}
		}
		auto issue_xin = [&](int k) {
			const int L = chunk_len(k), sl = k % KB_RV3_DI;
			const unsigned b = 4u * (unsigned)((L + 3) & ~3);            // (a ragged tail reads up to 3 floats past n, inside the row: stride % 4 == 0)
			kb_mbar_expect_tx(&S.bar_xin[sl], b);
			kb_bulk_g2s(&S.xin[sl][0], X + (size_t)k * Lc, b, &S.bar_xin[sl]);
		};
		for (int k = 0; k < KB_RV3_DI && k < K; k++) issue_xin(k);
		for (int k = KB_RV3_DI; k < K; k++) {
			kb_wait_ge(&S.w_done, k - KB_RV3_DI + 2);                // W reads an io chunk BEHIND its signal: chunk k - DI is done with once chunk k - DI + 1 is signalled
			issue_xin(k);
		}
		return;
	}

	if (role == KB_RV3_E) {
		// ================================================================ E: early cascade in >> lpf >> hpf (Reverb.k:87), lane 0 = LPF(chunk s), lane 1 = HPF(chunk s - 1)
		float z0 = 0.f, z1 = 0.f, b0 = 0.f, b1 = 0.f, b2 = 0.f, a1 = 0.f, a2 = 0.f;
		if (lane < 2) {
			const KbBiquad& f = lane == 0 ? rv.lpf[side] : rv.hpf[side];
			z0 = f.z0; z1 = f.z1; b0 = f.b0; b1 = f.b1; b2 = f.b2; a1 = f.a1; a2 = f.a2;
		}
		for (int s = 0; s <= K; s++) {
			if (s < K) kb_mbar_wait(&S.bar_xin[s % KB_RV3_DI], (unsigned)(s / KB_RV3_DI) & 1u);
			if (s >= 1) kb_wait_ge(&S.t_done, s - KB_RV3_DE);        // xf slot (s - 1) % DE was last read by T(s - 1 - DE)
			const int k = s - lane;
			if (lane == 0) KB_RV3_TR(KB_RV3_E, s, 0);
			if (lane < 2 && k >= 0 && k < K) {
				const float* xi = lane == 0 ? S.xin[k % KB_RV3_DI] : S.ylp[k & 1];
				float* xo = lane == 0 ? S.ylp[k & 1] : S.xf[k % KB_RV3_DE];
				kb_rv2_filter_row(xi, xo, chunk_len(k), b0, b1, b2, a1, a2, z0, z1);
			}
			__syncwarp();
			if (lane == 0) KB_RV3_TR(KB_RV3_E, s, 1);
			if (lane == 0) kb_signal(&S.e_done, s + 1);
		}
		if (lane == 0) { rv.lpf[side].z0 = z0; rv.lpf[side].z1 = z1; }
		if (lane == 1) { rv.hpf[side].z0 = z0; rv.hpf[side].z1 = z1; }
		return;
	}

	if (role == KB_RV3_T) {
		// ================================================================ T: early ring write and the taps (Reverb.k:86-92), thread = frame
		const int tt = rslot * 32 + lane;
		const int count = rv.count;
		kb_mbar_wait(&S.bar_res, 0u);
		int ebase = epos0;
		for (int k = 0; k < K; k++) {
			const int L = chunk_len(k);
			if (rslot == 0) { kb_wait_ge(&S.e_done, k + 2); kb_wait_ge(&S.w_done, k - KB_RV3_DR + 1); }    // xf(k) complete; r1 slot k % DR was read by W(k - DR)
			kb_bar_group(2, 96);
			if (tt == 0) KB_RV3_TR(KB_RV3_T, k, 0);
			if (tt < L) {
				int idx = ebase + tt; if (idx >= KB_RV3_ESIZE) idx -= KB_RV3_ESIZE;
				const float v = S.xf[k % KB_RV3_DE][tt];
				S.er[idx] = v; ringe[idx] = v;
				// (taps never reach into this chunk: Lc <= shortest tap - 3; older samples were published by the barrier of their chunk)
				int pos = idx + 1; if (pos >= KB_RV3_ESIZE) pos -= KB_RV3_ESIZE;      // position after this frame's write
				const float posf = (float)(pos - 1);
				float acc = 0.f;
				// batches of 5 taps, branch-free: a tap index past `count` is clamped for the loads and only its accumulation is predicated off, so the
				// address chains (float subtract, truncation, two dependent shared-memory gathers) of a batch overlap
				#pragma unroll
				for (int d0 = 0; d0 < KB_RV_MAXREFL; d0 += 5) {
					float va[5], vb[5], fr[5], gn[5];
					#pragma unroll
					for (int j = 0; j < 5; j++) {
						const int d = min(d0 + j, count - 1);
						float read = posf - S.times[d]; if (read < 0.f) read += KB_RV3_ESIZE;      // Stereo::Delay::tap(float)  klang.h:4668-4681
						const float fl = floorf(read); fr[j] = read - fl;
						const int ii = (int)read, jj = (ii == KB_RV3_ESIZE - 1) ? 0 : ii + 1;
						va[j] = S.er[ii]; vb[j] = S.er[jj]; gn[j] = S.gg[d];
					}
					#pragma unroll
					for (int j = 0; j < 5; j++) {
						const float term = (va[j] * (1.f - fr[j]) + vb[j] * fr[j]) * gn[j];          // out += delay(times[d]) * gains[d]  Reverb.k:89-90
						if (d0 + j < count) acc += term;
					}
				}
				S.r1[k % KB_RV3_DR][tt] = acc;
			}
			ebase += L; if (ebase >= KB_RV3_ESIZE) ebase -= KB_RV3_ESIZE;
			kb_bar_group(2, 96);
			if (tt == 0) KB_RV3_TR(KB_RV3_T, k, 1);
			if (tt == 0) kb_signal(&S.t_done, k + 1);
		}
		if (tt == 0) ed.position = (int)(((long long)epos0 + n) % KB_RV3_ESIZE);
		return;
	}

	if (role == KB_RV3_P) {
		// ================================================================ P (exact mode): ring windows -> Delay::process interpolation -> pre-multiplied operands
		// thread pt owns ticks pt and pt + 96 of every line: all 32 gathers of a chunk are issued before the first of them is consumed
		const int pt = rslot * 32 + lane;
		kb_mbar_wait(&S.bar_res, 0u);
		int rb[8];                                                   // index of the read head's sample in each line's shared-memory ring
		#pragma unroll
		for (int l = 0; l < 8; l++) rb[l] = S.lro[l];
		for (int k = 0; k < K; k++) {
			const int ticks = 2 * chunk_len(k), sl = k & 1;
			kb_wait_ge_group(&S.w_done, k - 1, rslot == 0, 1, 96);   // chunks 0 .. k-2 written: the windows of chunk k are complete, xq slot k & 1 is free
			if (pt == 0) KB_RV3_TR(KB_RV3_P, k, 0);
			const int tk1 = pt + 96;
			const bool has1 = tk1 < ticks;
			float xa[8][2], xb[8][2];
			#pragma unroll
			for (int l = 0; l < 8; l++) {
				const float* ring = &S.lr[S.lbase[l]];
				const int cap = S.lcap[l];
				int i = rb[l] + pt; if (i >= cap) i -= cap;
				int j = i + 1; if (j >= cap) j -= cap;
				xa[l][0] = ring[i]; xb[l][0] = ring[j];
				i = rb[l] + (has1 ? tk1 : pt); if (i >= cap) i -= cap;
				j = i + 1; if (j >= cap) j -= cap;
				xa[l][1] = ring[i]; xb[l][1] = ring[j];
				rb[l] += ticks; if (rb[l] >= cap) rb[l] -= cap;
			}
			#pragma unroll
			for (int l = 0; l < 8; l++) {
				const float fr = S.frac[l], b0 = S.b0[l], b1 = S.b1[l], b2 = S.b2[l];
				const float x0 = xa[l][0] + fr * (xb[l][0] - xa[l][0]);          // Delay::process  klang.h:3461-3473
				const float x1 = xa[l][1] + fr * (xb[l][1] - xa[l][1]);
				if (pt < ticks) S.xq[sl][l][pt] = make_float4(b0 * x0, b1 * x0, b2 * x0, 0.f);
				if (has1) S.xq[sl][l][tk1] = make_float4(b0 * x1, b1 * x1, b2 * x1, 0.f);
			}
			kb_bar_group(1, 96);
			if (pt == 0) KB_RV3_TR(KB_RV3_P, k, 1);
			if (pt == 0) kb_signal(&S.p_done, k + 1);
		}
		return;
	}

	if (role == KB_RV3_F) {
		// ================================================================ F (exact mode): the 8 line filters, lane = line, strictly in order
		float z0 = 0.f, z1 = 0.f, a1 = 0.f, a2 = 0.f;
		if (lane < 8) { const KbBiquad& f = kb_rv_side_line(rv, side, lane).filter; z0 = f.z0; z1 = f.z1; a1 = f.a1; a2 = f.a2; }
		for (int k = 0; k < K; k++) {
			kb_wait_ge(&S.p_done, k + 1);
			if (lane == 0) KB_RV3_TR(KB_RV3_F, k, 0);
			if (lane < 8) kb_rv3_filter_row(S.xq[k & 1][lane], S.y[k & 1][lane], 2 * chunk_len(k), a1, a2, z0, z1);
			__syncwarp();
			if (lane == 0) KB_RV3_TR(KB_RV3_F, k, 1);
			if (lane == 0) kb_signal(&S.f_done, k + 1);
		}
		if (lane < 8) { KbBiquad& f = kb_rv_side_line(rv, side, lane).filter; f.z0 = z0; f.z1 = z1; }
		return;
	}

	if (role == KB_RV3_S) {
		// ================================================================ S (tolerance mode): one warp per line, parallel scan over the chunk's ticks
		const int l = rslot;
		KbRv3ScanCoef sc;
		kb_rv3_scan_coef(kb_rv_side_line(rv, side, l).filter, sc);
		float z0 = kb_rv_side_line(rv, side, l).filter.z0, z1 = kb_rv_side_line(rv, side, l).filter.z1;
		const float fr = S.frac[l];
		const float* ring = &S.lr[S.lbase[l]];
		const int cap = S.lcap[l];
		int rb = S.lro[l];
		kb_mbar_wait(&S.bar_res, 0u);
		for (int k = 0; k < K; k++) {
			const int sl = k & 1, ticks = 2 * chunk_len(k);
			kb_wait_ge_group(&S.w_done, k - 1, l == 0, 4, 256);      // the window of chunk k is complete, y slot k & 1 is free
			if (lane == 0) KB_RV3_TR(8 + l, k, 0);
			kb_rv3_scan_chunk(sc, ring, cap, rb, fr, ticks, S.y[sl][l], z0, z1, lane);
			rb += ticks; if (rb >= cap) rb -= cap;
			__syncwarp();
			if (lane == 0) KB_RV3_TR(8 + l, k, 1);
			if (lane == 0) asm volatile("red.release.cta.shared.add.s32 [%0], 1;" :: "r"(kb_smem_u32(&S.f_cnt[sl])) : "memory");   // chunk k is complete when its parity's count reaches 8 (k / 2 + 1)
		}
		if (lane == 0) { KbBiquad& f = kb_rv_side_line(rv, side, l).filter; f.z0 = z0; f.z1 = z1; }
		return;
	}

	// ==================================================================== W: FDN matrix, ring writes, mid -> late, output mix (Reverb.k:152-169, 212-231, 272)
	// thread = (frame t, stage): stage 0 = mid, 1 = late (5 warps: 2 x 80 threads).  Both stages of a frame are independent once the filter outputs
	// are known: late's input r2 is the in-order sum of mid's second-tick outputs, which the stage-1 thread forms itself (same expression, same bits).
	// The shared-memory ring writes — all that the next windows wait for — come first and are signalled at once; the write-through to the global
	// rings and the output mix follow behind the signal.
	{
		const int wt = rslot * 32 + lane;
		const int stage = wt >= KB_RV3_LMAX ? 1 : 0, t = wt - stage * KB_RV3_LMAX, base = stage * 4;
		const float dry = c[0].value, wet = side == 0 ? c[4].value : 0.f;        // Reverb.k:272 (Q7): the right wet gain is the literal 0
		const float cE = c[1].value, cM = c[2].value, cL = c[3].value;
		const float M[4][4] = { { 0, 1, 1, -1 }, { -1, 0, -1, 1 }, { -1, 1, 0, -1 }, { 1, -1, 1, 0 } };      // Reverb.k:158-161
		int ws[4], cap[4];                                           // the four lines of this thread's stage: write head in the shared-memory ring, its capacity
		float* sring[4]; float gain[4], gmid[4];
		#pragma unroll
		for (int q = 0; q < 4; q++) {
			const int l = base + q;
			cap[q] = S.lcap[l];
			ws[q] = S.lwo[l]; if (ws[q] >= cap[q]) ws[q] -= cap[q];
			sring[q] = &S.lr[S.lbase[l]];
			gain[q] = S.gain[l]; gmid[q] = S.gain[q];
		}
		int gp[8], gs[8];                                            // all 8 lines: write head in the global ring / the shared-memory ring (write-through)
		#pragma unroll
		for (int l = 0; l < 8; l++) { gp[l] = S.wpos0[l]; gs[l] = S.lwo[l]; if (gs[l] >= S.lcap[l]) gs[l] -= S.lcap[l]; }
		kb_mbar_wait(&S.bar_res, 0u);                                // (the resident spans must have landed before this role writes behind them)
		int cpar = 0;
		for (int k = 0; k < K; k++, cpar ^= 1) {
			const int L = chunk_len(k), sl = k & 1;
			if (rslot == 0) {
				if (MODE == 1) kb_wait_ge(&S.f_cnt[sl], 8 * ((k >> 1) + 1)); else kb_wait_ge(&S.f_done, k + 1);
				kb_wait_ge(&S.t_done, k + 1);
			}
			kb_bar_group(3, 160);
			if (wt == 0) KB_RV3_TR(KB_RV3_W, k, 0);
			float fb[4], r1 = 0.f, r2 = 0.f, sum = 0.f;
			const bool on = t < L;
			if (on) {
				r1 = S.r1[k % KB_RV3_DR][t];
				float dv[4], sv[4];
				#pragma unroll
				for (int j = 0; j < 4; j++) {
					const float2 yy = *reinterpret_cast<const float2*>(&S.y[sl][base + j][2 * t]);
					dv[j] = yy.x * gain[j];                                      // FilteredDelay::process  Reverb.k:130-132
					sv[j] = yy.y * gain[j];
				}
				sum = sv[0];
				sum = sum + sv[1];
				sum = sum + sv[2];
				sum = sum + sv[3];
				float in = r1;
				if (stage == 1) {                                            // late's input = mid's output of this frame
					r2 = S.y[sl][0][2 * t + 1] * gmid[0];
					r2 = r2 + S.y[sl][1][2 * t + 1] * gmid[1];
					r2 = r2 + S.y[sl][2][2 * t + 1] * gmid[2];
					r2 = r2 + S.y[sl][3][2 * t + 1] * gmid[3];
					in = r2;
				}
				#pragma unroll
				for (int q = 0; q < 4; q++) {
					// feedback * delays + in, row q with its literal 0 / +-1 products (Reverb.k:158-163, klang.h:1446-1470)
					fb[q] = (M[q][0] * dv[0] + M[q][1] * dv[1] + M[q][2] * dv[2] + M[q][3] * dv[3]) + in;
					int s0 = ws[q] + 2 * t; if (s0 >= cap[q]) s0 -= cap[q];
					int sa = s0 + 1; if (sa >= cap[q]) sa -= cap[q];
					sring[q][sa] = fb[q];                                        // second tick of this frame writes fb
					if (t + 1 < L) { int sb = s0 + 2; if (sb >= cap[q]) sb -= cap[q]; sring[q][sb] = fb[q]; }   // = first tick of the next frame
					else S.carry[cpar ^ 1][base + q] = fb[q];
					if (t == 0) sring[q][s0] = S.carry[cpar][base + q];
				}
			}
			kb_bar_group(3, 160);
			if (wt == 0) { KB_RV3_TR(KB_RV3_W, k, 1); kb_signal(&S.w_done, k + 1); }
			// ---- behind the signal: write-through of the chunk's 2 L new samples of every line from the shared-memory rings to the global rings
			// (read again only by a later launch), thread = position: full-sector coalesced stores; and the output
			#pragma unroll
			for (int l = 0; l < 8; l++) {
				if (wt < 2 * L) {
					int g = gp[l] + wt; if (g >= S.lsize[l]) g -= S.lsize[l];
					int sp = gs[l] + wt; if (sp >= S.lcap[l]) sp -= S.lcap[l];
					(rings + S.lring[l])[g] = S.lr[S.lbase[l] + sp];
				}
				gp[l] += 2 * L; if (gp[l] >= S.lsize[l]) gp[l] -= S.lsize[l];
				gs[l] += 2 * L; if (gs[l] >= S.lcap[l]) gs[l] -= S.lcap[l];
			}
			if (on && stage == 1) {
				const float refl = (r1 * cE + r2 * cM) + sum * cL;
				X[(size_t)k * Lc + t] = S.xin[k % KB_RV3_DI][t] * dry + refl * wet;      // Reverb.k:272
			}
			#pragma unroll
			for (int q = 0; q < 4; q++) { ws[q] += 2 * L; if (ws[q] >= cap[q]) ws[q] -= cap[q]; }
		}
		kb_bar_group(3, 160);
		if (wt < 8) {
			KbRvFDelay& d = kb_rv_side_line(rv, side, wt);
			d.in = S.carry[cpar][wt];
			d.delay.position = (int)(((long long)d.delay.position + 2LL * n) % d.delay.SIZE);
			d.delay.last_position = (int)(((long long)d.delay.last_position + 2LL * n) % d.delay.SIZE);
		}
	}
}
#endif  // __CUDACC__

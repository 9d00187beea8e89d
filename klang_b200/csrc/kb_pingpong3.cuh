// klang-b200 — PingPong.k (examples/PingPong.k:36-71), fused schedule: ONE launch per block, one CTA per instance.
//
// The delay network of PingPong.k is time-parallel once its control smoothers have settled (kb_fx_parallel.cuh: the read heads then keep
// a constant distance to the write heads); what stays serial is the DC blocker of each channel, Biquad::HPF(50 Hz, Q = 1), whose fp32
// rounding noise is 130x the parity bar (profiles/r02_reverb_tolerance.txt) — it must run in the reference's order.  The round-1 schedule
// ran plan, network (whole block), filter (whole block, one thread), store, finish as separate phases and launches: 70 us per 4096-frame
// block against the 35 us the filter chain needs by itself.  Here:
//   * thread 0 classifies the instance in the kernel (the same bit-exact fixed-point test as kb_pingpong_plan_kernel) — no plan / finish
//     launches; an instance that does not qualify is evaluated frame by frame by that thread (same code as kb_fx_seq_kernel);
//   * four producer warps evaluate the delay network in chunks of 128 frames, eight chunks ahead of the filter, and hand it PRE-MULTIPLIED
//     operands (b0 x, b1 x, b2 x): the filter warp (lane 0 = left channel, lane 1 = right) executes only the recurrence
//     y = p0 + z0; z0 = (p1 - a1 y) + z1; z1 = p2 - a2 y (kb_rv3_filter_row) — the roundings of Biquad::Filter::process, klang.h:5605-5612;
//   * the producers store the filtered chunks behind the filter; hand-over by progress counters in shared memory, no CTA-wide barrier.
#pragma once
#include "kb_reverb3.cuh"

#define KB_PP3_CH 128                         // frames per chunk
#define KB_PP3_NSLOT 8                        // chunks in flight
#define KB_PP3_NT 160                         // warp 0 = filter, warps 1-4 = producers / storers

struct KbPp3Smem {
	float4 xq[KB_PP3_NSLOT][2][KB_PP3_CH + 4];   // filter operands per (chunk slot, channel, frame)
	float y[KB_PP3_NSLOT][2][KB_PP3_CH];         // filtered samples
	KbFxPlan plan;
	int p_done, f_done;
};

__global__ void __launch_bounds__(KB_PP3_NT) kb_pingpong3_kernel(KbFxHdr* __restrict__ hdrs, KbPingPong* __restrict__ states, KbFxPlan* __restrict__ plan_out,
                                                                 float* __restrict__ rings, float* __restrict__ io, int n, int stride, KbFs fs) {
	__shared__ __align__(16) KbPp3Smem S;
	const int inst = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
	KbPingPong& p = states[inst];
	float* L = io + (size_t)inst * 2 * stride; float* R = L + stride;
	if (tid == 0) {
		// the control half of the frame (smoothers, LFO, delay hand-over) executed once on a copy: parallel iff it is at a bit-exact fixed point
		KbFxHdr h = hdrs[inst];
		KbPingPong q = p;
		const KbFxHdr& h0 = hdrs[inst];
		KbFxPlan pl;
		kb_pingpong_control(fs, h, q, pl.gain, pl.delay, pl.dry);
		bool fixed = kb_same_bits(q.delay, p.delay) && kb_same_bits(q.lfo.position, p.lfo.position) && kb_same_bits(q.lfo.increment, p.lfo.increment) &&
		             kb_same_bits(q.lfo.frequency, p.lfo.frequency);
		for (int c = 0; c < 6; c++) fixed = fixed && kb_same_bits(h.controls[c].value, h0.controls[c].value) && kb_same_bits(h.controls[c].smoothed, h0.controls[c].smoothed);
		const float dl = pl.delay * fs.f, dr = 0.5f * pl.delay * fs.f;
		pl.chunk = (int)fminf(dl, dr) - 4;
		// (far end: the block's later frames must not overwrite ring slots its earlier frames still read)
		pl.mode = (fixed && pl.chunk >= n && (float)n + dl + 4.f < (float)p.left.SIZE) ? KB_PLAN_PARALLEL : KB_PLAN_SEQUENTIAL;
		S.plan = pl; plan_out[inst] = pl;
		S.p_done = 0; S.f_done = 0;
		if (pl.mode == KB_PLAN_SEQUENTIAL) {
			// frame by frame (Stereo::Effect::process, klang.h:4708-4716): exact for any control motion
			KbFxHdr hs = hdrs[inst];
			KbPingPong s = p;
			for (int i = 0; i < n; i++) { float ol, orr; kb_pingpong_frame(fs, hs, s, rings, L[i], R[i], ol, orr); L[i] = ol; R[i] = orr; }
			hdrs[inst] = hs; p = s;
		}
	}
	__syncthreads();
	const KbFxPlan pl = S.plan;
	if (pl.mode != KB_PLAN_PARALLEL) return;
	const int K = (n + KB_PP3_CH - 1) / KB_PP3_CH;
	auto chunk_len = [&](int c) { return min(KB_PP3_CH, n - c * KB_PP3_CH); };

	if (warp == 0) {
		// ---- the two DC blockers, strictly in order
		float z0 = 0.f, z1 = 0.f, a1 = 0.f, a2 = 0.f;
		if (lane < 2) { const KbBiquad& b = p.dc[lane]; z0 = b.z0; z1 = b.z1; a1 = b.a1; a2 = b.a2; }
		for (int c = 0; c < K; c++) {
			kb_wait_ge(&S.p_done, c + 1);
			if (lane < 2) kb_rv3_filter_row(S.xq[c % KB_PP3_NSLOT][lane], S.y[c % KB_PP3_NSLOT][lane], chunk_len(c), a1, a2, z0, z1);
			__syncwarp();
			if (lane == 0) kb_signal(&S.f_done, c + 1);
		}
		if (lane < 2) { p.dc[lane].z0 = z0; p.dc[lane].z1 = z1; }
		return;
	}

	// ---- producers / storers: thread = frame of the chunk
	const int t = tid - 32;
	float* ringl = rings + p.left.ring; float* ringr = rings + p.right.ring;
	const int SIZE = p.left.SIZE;
	const float gain = pl.gain, dry = pl.dry;
	const float dl = pl.delay * fs.f, dr = 0.5f * pl.delay * fs.f;            // left.set(delay*fs), right.set(0.5f*delay*fs)   PingPong.k:63-64
	const int pl0 = p.left.position, pr0 = p.right.position;
	const float lb0 = p.dc[0].b0, lb1 = p.dc[0].b1, lb2 = p.dc[0].b2, rb0 = p.dc[1].b0, rb1 = p.dc[1].b1, rb2 = p.dc[1].b2;
	auto store_chunk = [&](int c) {
		const int len = chunk_len(c), sl = c % KB_PP3_NSLOT;
		if (t < len) { L[c * KB_PP3_CH + t] = S.y[sl][0][t]; R[c * KB_PP3_CH + t] = S.y[sl][1][t]; }
	};
	for (int c = 0; c < K; c++) {
		const int len = chunk_len(c), sl = c % KB_PP3_NSLOT, f = c * KB_PP3_CH + t;
		float inl = 0.f, inr = 0.f, ra = 0.f, rb = 0.f, rc = 0.f, la = 0.f, lb = 0.f, lc = 0.f, fl = 0.f, fr = 0.f;
		int posl = 0, posr = 0;
		if (t < len) {
			posl = (int)(((long long)pl0 + f) % SIZE); posr = (int)(((long long)pr0 + f) % SIZE);
			// Delay::set (klang.h:3480-3489) relative to the write heads of this frame; each line is read-ticked twice per frame (Q13): the two ticks
			// of a line interpolate ring[i], ring[i+1] and ring[i+1], ring[i+2] with the same fraction
			float rl = (float)(posl - 1) - dl; if (rl < 0.f) rl += SIZE;
			const int il = (int)rl; fl = rl - il;
			float rr = (float)(posr - 1) - dr; if (rr < 0.f) rr += SIZE;
			const int ir = (int)rr; fr = rr - ir;
			const int ir1 = (ir + 1) % SIZE, ir2 = (ir1 + 1) % SIZE, il1 = (il + 1) % SIZE, il2 = (il1 + 1) % SIZE;
			inl = L[f]; inr = R[f];
			ra = ringr[ir]; rb = ringr[ir1]; rc = ringr[ir2];
			la = ringl[il]; lb = ringl[il1]; lc = ringl[il2];
			if (f == n - 1) {   // read-head state after the block: two ticks past the last set()
				p.left.last_position = (il2) % SIZE; p.left.last_fraction = fl; p.left.time = dl;
				p.right.last_position = (ir2) % SIZE; p.right.last_fraction = fr; p.right.time = dr;
			}
		}
		if (c >= KB_PP3_NSLOT) { kb_wait_ge_group(&S.f_done, c - KB_PP3_NSLOT + 1, warp == 1, 2, 128); store_chunk(c - KB_PP3_NSLOT); }     // the slot's previous chunk is filtered: store it, the slot is free
		if (t < len) {
			const float rtick = ra + fr * (rb - ra);                                // right's first read tick          PingPong.k:66
			ringl[posl] = inl + rtick * gain;                                        // left write
			const float ltick = la + fl * (lb - la);                                // left's first read tick
			const float prel = dry * inl + (1.f - dry) * ltick;
			const float ltick2 = lb + fl * (lc - lb);                               // left's second read tick         PingPong.k:67
			ringr[posr] = inr + ltick2 * gain;                                       // right write
			const float rtick2 = rb + fr * (rc - rb);                               // right's second read tick
			const float prer = dry * inr + (1.f - dry) * rtick2;
			S.xq[sl][0][t] = make_float4(lb0 * prel, lb1 * prel, lb2 * prel, 0.f);
			S.xq[sl][1][t] = make_float4(rb0 * prer, rb1 * prer, rb2 * prer, 0.f);
			if (f == n - 1) { p.left.out = ltick2; p.right.out = rtick2; }
		}
		kb_bar_group(1, 128);
		if (t == 0) kb_signal(&S.p_done, c + 1);
	}
	for (int c = max(0, K - KB_PP3_NSLOT); c < K; c++) { kb_wait_ge_group(&S.f_done, c + 1, warp == 1, 2, 128); store_chunk(c); }
	if (t == 0) {
		p.left.position = (int)(((long long)pl0 + n) % SIZE);
		p.right.position = (int)(((long long)pr0 + n) % SIZE);
	}
}

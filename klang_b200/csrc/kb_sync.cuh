// klang-b200 — hand-over between the warp roles of one CTA without a CTA-wide barrier (used by kb_reverb3.cuh, kb_pingpong3.cuh, kb_tiled.cuh).
#pragma once
#include "kb_math.cuh"

// named barrier among the `threads` threads of one role
KB_D void kb_bar_group(int id, int threads) { asm volatile("bar.sync %0, %1;" :: "r"(id), "r"(threads) : "memory"); }

// ---- hand-over primitives: a progress counter in shared memory, written with st.release.cta by ONE thread of the producing role (after the
// role's own barrier, which orders the other threads' writes before it) and polled with ld.acquire.cta.  No sequentially-consistent fence
// (__threadfence_block() is MEMBAR.SC.CTA, which also waits for the thread's outstanding global stores: ~1000 cycles per hand-over here).
KB_D int kb_ld_acquire(const int* p) {
	int v;
	asm volatile("ld.acquire.cta.shared.b32 %0, [%1];" : "=r"(v) : "r"((unsigned)__cvta_generic_to_shared(p)) : "memory");
	return v;
}
KB_D void kb_wait_ge(const int* counter, int target) {
	while (kb_ld_acquire(counter) < target) __nanosleep(48);
}
KB_D void kb_signal(int* counter, int value) {
	asm volatile("st.release.cta.shared.b32 [%0], %1;" :: "r"((unsigned)__cvta_generic_to_shared(counter)), "r"(value) : "memory");
}
// a multi-warp role waits: its first warp polls, the others sleep at the role's named barrier (no issue slots, no shared-memory polling)
KB_D void kb_wait_ge_group(const int* counter, int target, bool first_warp, int bar_id, int threads) {
	if (first_warp) kb_wait_ge(counter, target);
	asm volatile("bar.sync %0, %1;" :: "r"(bar_id), "r"(threads) : "memory");
}

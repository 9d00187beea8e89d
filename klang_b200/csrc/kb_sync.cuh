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
// (measurement variant) the same hand-over with a plain volatile store: the producing role's barrier has already ordered its shared-memory
// writes, and shared memory has no cache to leave them in; no MEMBAR in front of the store
KB_D void kb_signal_v(bool relaxed, int* counter, int value) {
	if (relaxed) asm volatile("st.volatile.shared.b32 [%0], %1;" :: "r"((unsigned)__cvta_generic_to_shared(counter)), "r"(value) : "memory");
	else kb_signal(counter, value);
}
// a multi-warp role waits: its first warp polls, the others sleep at the role's named barrier (no issue slots, no shared-memory polling)
KB_D void kb_wait_ge_group(const int* counter, int target, bool first_warp, int bar_id, int threads) {
	if (first_warp) kb_wait_ge(counter, target);
	asm volatile("bar.sync %0, %1;" :: "r"(bar_id), "r"(threads) : "memory");
}

// ---- mbarrier objects in shared memory (SASS SYNCS): hardware-parked waits instead of polling, one arrival per producing warp
KB_D unsigned kb_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
KB_D void kb_mbar_init(unsigned long long* bar, unsigned count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(kb_smem_u32(bar)), "r"(count) : "memory"); }
KB_D void kb_mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(kb_smem_u32(bar)), "r"(bytes) : "memory");
}
KB_D void kb_mbar_wait(unsigned long long* bar, unsigned parity) {
	asm volatile(
		"{\n\t.reg .pred p;\n\t"
		"KB_MBAR_WAIT_%=:\n\t"
		"mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
		"@p bra KB_MBAR_DONE_%=;\n\t"
		"bra KB_MBAR_WAIT_%=;\n\t"
		"KB_MBAR_DONE_%=:\n\t}"
		:: "r"(kb_smem_u32(bar)), "r"(parity) : "memory");
}
// one arrival (release at CTA scope: the arriving thread's earlier shared-memory writes — and, behind a __syncwarp(), its warp's — are
// visible to a thread whose wait on this phase has returned)
KB_D void kb_mbar_arrive(unsigned long long* bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(kb_smem_u32(bar)) : "memory"); }
// non-blocking look at a phase (acquire when it has completed)
KB_D bool kb_mbar_test(unsigned long long* bar, unsigned parity) {
	unsigned ok;
	asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(kb_smem_u32(bar)), "r"(parity) : "memory");
	return ok != 0;
}

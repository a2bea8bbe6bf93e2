// warp_emu.h -- test infrastructure: runs warp-cooperative device code on the host, 32 lanes as 32 coroutines.
//
// Every *_sync intrinsic is a rendezvous: a lane deposits its operand, yields, and is resumed once all 32 lanes of
// the warp have deposited theirs (double-buffered, so a lane that runs ahead to its next rendezvous cannot clobber
// operands others still have to read).  A lane that returns while others wait at a rendezvous is a divergence bug
// and aborts -- full-mask intrinsics under divergent control flow are undefined on the GPU, the emulator makes
// them loud.  Only the intrinsics the repo's warp code uses are provided.
#pragma once
#include <ucontext.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <functional>
#include <vector>

namespace warp_emu {

struct Warp {
    ucontext_t sched;
    ucontext_t ctx[32];
    std::vector<char> stack[32];
    uint64_t buf[2][32];
    uint32_t gen[32];
    bool done[32], waiting[32];
    int cur;
    std::function<void()> body;
};

inline Warp *&current() {
    static thread_local Warp *w = nullptr;
    return w;
}
inline int lane_id() { return current()->cur; }

inline void trampoline() {
    Warp *w = current();
    const int me = w->cur;
    w->body();
    w->done[me] = true;
    swapcontext(&w->ctx[me], &w->sched);
}

// deposit v, wait for the other lanes, return the 32 operands of this rendezvous
inline const uint64_t *exchange(uint64_t v) {
    Warp *w = current();
    const int me = w->cur;
    const uint32_t g = w->gen[me]++;
    w->buf[g & 1][me] = v;
    w->waiting[me] = true;
    swapcontext(&w->ctx[me], &w->sched);
    w->cur = me;
    return w->buf[g & 1];
}

// run body() once per lane in lockstep
inline void run_warp(const std::function<void()> &body) {
    Warp *w = new Warp();
    Warp *outer = current();
    current() = w;
    w->body = body;
    for (int l = 0; l < 32; l++) {
        w->stack[l].resize(256 << 10);
        w->gen[l] = 0;
        w->done[l] = w->waiting[l] = false;
        getcontext(&w->ctx[l]);
        w->ctx[l].uc_stack.ss_sp = w->stack[l].data();
        w->ctx[l].uc_stack.ss_size = w->stack[l].size();
        w->ctx[l].uc_link = &w->sched;
        makecontext(&w->ctx[l], (void (*)()) trampoline, 0);
    }
    for (;;) {
        int n_done = 0, n_wait = 0;
        for (int l = 0; l < 32; l++) {
            if (w->done[l]) { n_done++; continue; }
            w->waiting[l] = false;
            w->cur = l;
            swapcontext(&w->sched, &w->ctx[l]);
            if (w->done[l]) n_done++; else n_wait++;
        }
        if (n_wait == 0) break;
        if (n_done != 0) {
            fprintf(stderr, "warp_emu: %d lane(s) returned while %d wait at a full-mask intrinsic; rendezvous passed per lane:", n_done, n_wait);
            for (int l = 0; l < 32; l++) fprintf(stderr, " %u%s", w->gen[l], w->done[l] ? "r" : "");
            fprintf(stderr, "\n");
            abort();
        }
        // all 32 wait at (what must be) the same rendezvous
        for (int l = 1; l < 32; l++)
            if (w->gen[l] != w->gen[0]) { fprintf(stderr, "warp_emu: lanes at different rendezvous\n"); abort(); }
    }
    current() = outer;
    delete w;
}

template <class T> inline uint64_t to_bits(T v) {
    uint64_t b = 0;
    static_assert(sizeof(T) <= 8, "operand");
    __builtin_memcpy(&b, &v, sizeof(T));
    return b;
}
template <class T> inline T from_bits(uint64_t b) {
    T v;
    __builtin_memcpy(&v, &b, sizeof(T));
    return v;
}

}  // namespace warp_emu

// ---- the intrinsics (full mask only) ------------------------------------------------------------------------
template <class T> inline T __shfl_sync(unsigned, T v, int src) {
    const uint64_t *x = warp_emu::exchange(warp_emu::to_bits(v));
    return warp_emu::from_bits<T>(x[src & 31]);
}
template <class T> inline T __shfl_up_sync(unsigned, T v, int d) {
    const int me = warp_emu::lane_id();
    const uint64_t *x = warp_emu::exchange(warp_emu::to_bits(v));
    return warp_emu::from_bits<T>(x[me - d >= 0 ? me - d : me]);
}
template <class T> inline T __shfl_down_sync(unsigned, T v, int d) {
    const int me = warp_emu::lane_id();
    const uint64_t *x = warp_emu::exchange(warp_emu::to_bits(v));
    return warp_emu::from_bits<T>(x[me + d < 32 ? me + d : me]);
}
template <class T> inline T __shfl_xor_sync(unsigned, T v, int m) {
    const int me = warp_emu::lane_id();
    const uint64_t *x = warp_emu::exchange(warp_emu::to_bits(v));
    return warp_emu::from_bits<T>(x[(me ^ m) & 31]);
}
inline unsigned __ballot_sync(unsigned, bool p) {
    const uint64_t *x = warp_emu::exchange(p ? 1 : 0);
    unsigned m = 0;
    for (int l = 0; l < 32; l++) m |= (unsigned) (x[l] & 1) << l;
    return m;
}
inline bool __any_sync(unsigned m, bool p) { return __ballot_sync(m, p) != 0; }
inline bool __all_sync(unsigned m, bool p) { return __ballot_sync(m, p) == 0xffffffffu; }
inline int __reduce_min_sync(unsigned, int v) {
    const uint64_t *x = warp_emu::exchange(warp_emu::to_bits(v));
    int r = warp_emu::from_bits<int>(x[0]);
    for (int l = 1; l < 32; l++) { const int t = warp_emu::from_bits<int>(x[l]); if (t < r) r = t; }
    return r;
}
inline int __reduce_max_sync(unsigned, int v) {
    const uint64_t *x = warp_emu::exchange(warp_emu::to_bits(v));
    int r = warp_emu::from_bits<int>(x[0]);
    for (int l = 1; l < 32; l++) { const int t = warp_emu::from_bits<int>(x[l]); if (t > r) r = t; }
    return r;
}
inline int __reduce_add_sync(unsigned, int v) {
    const uint64_t *x = warp_emu::exchange(warp_emu::to_bits(v));
    int r = 0;
    for (int l = 0; l < 32; l++) r += warp_emu::from_bits<int>(x[l]);
    return r;
}
inline unsigned __reduce_or_sync(unsigned, unsigned v) {
    const uint64_t *x = warp_emu::exchange(v);
    unsigned r = 0;
    for (int l = 0; l < 32; l++) r |= (unsigned) x[l];
    return r;
}
inline void __syncwarp(unsigned = 0xffffffffu) { warp_emu::exchange(0); }
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __clz(int v) { return v ? __builtin_clz((unsigned) v) : 32; }
inline int __ffs(int v) { return __builtin_ffs(v); }
// position of the n-th (1-based) set bit of mask at or above base, 0xffffffff when there is none (offset > 0 form)
inline unsigned __fns(unsigned mask, unsigned base, int offset) {
    if (offset <= 0) return 0xffffffffu;  // (the other forms are not used here)
    for (unsigned b = base; b < 32; b++)
        if ((mask >> b) & 1u) { if (--offset == 0) return b; }
    return 0xffffffffu;
}

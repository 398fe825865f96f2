package lwse

/*
#include <stdlib.h>
#include <string.h>
#include "lwse.h"
*/
import "C"

import "unsafe"

// Arena hands out record tables in C memory.  cgo forbids C code to keep pointers into the Go
// heap, and DMA wants memory the collector never moves, so every table the engine sees is
// C-allocated: posix_memalign'd for the stateless entry points (Sweep, Place, GroupKeys), or a
// slice of the engine's own pinned, mapped patch arena (PatchArena) for the resident tick — the GPU
// reads those bytes in place, no staging copy.
//
// The same pattern, in C, is exercised by tests/host_c/cgo_pattern_check.c (this image has no Go
// toolchain): C-malloc'd tables, several threads on one handle.
type Arena struct{ blocks []unsafe.Pointer }

// Alloc returns n zeroed bytes, 64-byte aligned (every table base must be 16-byte aligned).
func (a *Arena) Alloc(n uintptr) unsafe.Pointer {
	var p unsafe.Pointer
	if C.posix_memalign(&p, 64, C.size_t(n+64)) != 0 {
		return nil
	}
	C.memset(p, 0, C.size_t(n+64))
	a.blocks = append(a.blocks, p)
	return p
}

// Free releases every block of the arena.
func (a *Arena) Free() {
	for _, p := range a.blocks {
		C.free(p)
	}
	a.blocks = nil
}

// LwsRecs / GroupRecs / PodStates / PodIdents view n rows of a block as a Go slice (no copy).
func LwsRecs(p unsafe.Pointer, n int) []C.lwse_lws_rec     { return unsafe.Slice((*C.lwse_lws_rec)(p), n) }
func GroupRecs(p unsafe.Pointer, n int) []C.lwse_group_rec { return unsafe.Slice((*C.lwse_group_rec)(p), n) }
func PodStates(p unsafe.Pointer, n int) []C.lwse_pod_state { return unsafe.Slice((*C.lwse_pod_state)(p), n) }
func PodIdents(p unsafe.Pointer, n int) []C.lwse_pod_ident { return unsafe.Slice((*C.lwse_pod_ident)(p), n) }

// PatchArena is the engine's pinned, mapped patch arena (lwse_resident_arena): row numbers and
// values written here are read by the scatter kernel over PCIe where they lie.
func (e *Engine) PatchArena(minBytes uint64) (base unsafe.Pointer, size uint64, err error) {
	var b unsafe.Pointer
	var n C.uint64_t
	err = e.check(C.lwse_resident_arena(e.h, C.uint64_t(minBytes), &b, &n))
	return b, uint64(n), err
}

// ResidentLoad makes the four input tables resident on the device.
func (e *Engine) ResidentLoad(t *LwsTables) error {
	ct := t.c()
	return e.check(C.lwse_resident_load(e.h, &ct))
}

// ResidentPlaceLoad makes the placement request table resident (rows grouped by namespace get
// the namespace-parallel kernels).
func (e *Engine) ResidentPlaceLoad(reqs *C.lwse_place_req, n, nNamespaces uint32) error {
	return e.check(C.lwse_resident_place_load(e.h, reqs, C.uint32_t(n), C.uint32_t(nNamespaces)))
}

// Tick is one pass of the work queues: watch-event patches in, changed result rows out.  The
// returned slices alias engine-owned pinned memory and are valid until the next resident call.
type TickResult struct {
	LwsRows   []uint32
	LwsOut    []C.lwse_lws_out
	GroupRows []uint32
	GroupOut  []C.lwse_group_out
	PlaceRows []uint32
	PlaceOut  []C.lwse_place_out
	Rounds    uint32
}

func (e *Engine) Tick(segs []C.lwse_patch_seg, flags uint32) (TickResult, error) {
	var t C.lwse_tick
	if len(segs) > 0 {
		// the segment descriptors themselves are only read during the call: Go memory is fine
		t.segs = (*C.lwse_patch_seg)(unsafe.Pointer(&segs[0]))
		t.n_segs = C.uint32_t(len(segs))
	}
	t.flags = C.uint32_t(flags)
	if err := e.check(C.lwse_resident_tick(e.h, &t)); err != nil {
		return TickResult{}, err
	}
	return TickResult{
		LwsRows:   unsafe.Slice((*uint32)(unsafe.Pointer(t.lws_rows)), int(t.n_lws)),
		LwsOut:    unsafe.Slice((*C.lwse_lws_out)(unsafe.Pointer(t.lws_out)), int(t.n_lws)),
		GroupRows: unsafe.Slice((*uint32)(unsafe.Pointer(t.group_rows)), int(t.n_groups)),
		GroupOut:  unsafe.Slice((*C.lwse_group_out)(unsafe.Pointer(t.group_out)), int(t.n_groups)),
		PlaceRows: unsafe.Slice((*uint32)(unsafe.Pointer(t.place_rows)), int(t.n_place)),
		PlaceOut:  unsafe.Slice((*C.lwse_place_out)(unsafe.Pointer(t.place_out)), int(t.n_place)),
		Rounds:    uint32(t.place_rounds),
	}, nil
}

// Submit enqueues a tick and returns at once; Wait blocks for the OLDEST submitted tick.  At most
// two ticks are in flight: the sweeper goroutine writes the next batch of watch-event patches into
// the other half of the arena and submits it while the GPU still works on the previous one.
// The patch bytes of a submitted tick must stay untouched until its Wait returned.
func (e *Engine) Submit(segs []C.lwse_patch_seg, flags uint32) error {
	var t C.lwse_tick
	if len(segs) > 0 {
		t.segs = (*C.lwse_patch_seg)(unsafe.Pointer(&segs[0]))
		t.n_segs = C.uint32_t(len(segs))
	}
	t.flags = C.uint32_t(flags)
	return e.check(C.lwse_resident_tick_submit(e.h, &t))
}

func (e *Engine) Wait() (TickResult, error) {
	var t C.lwse_tick
	if err := e.check(C.lwse_resident_tick_wait(e.h, &t)); err != nil {
		return TickResult{}, err
	}
	return tickResult(&t), nil
}

func tickResult(t *C.lwse_tick) TickResult {
	return TickResult{
		LwsRows:   unsafe.Slice((*uint32)(unsafe.Pointer(t.lws_rows)), int(t.n_lws)),
		LwsOut:    unsafe.Slice((*C.lwse_lws_out)(unsafe.Pointer(t.lws_out)), int(t.n_lws)),
		GroupRows: unsafe.Slice((*uint32)(unsafe.Pointer(t.group_rows)), int(t.n_groups)),
		GroupOut:  unsafe.Slice((*C.lwse_group_out)(unsafe.Pointer(t.group_out)), int(t.n_groups)),
		PlaceRows: unsafe.Slice((*uint32)(unsafe.Pointer(t.place_rows)), int(t.n_place)),
		PlaceOut:  unsafe.Slice((*C.lwse_place_out)(unsafe.Pointer(t.place_out)), int(t.n_place)),
		Rounds:    uint32(t.place_rounds),
	}
}

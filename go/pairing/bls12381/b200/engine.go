//go:build b200

// Package b200 is a BLS12-381 backend for kyber whose group arithmetic runs on an NVIDIA B200 through
// libb2kyber.so.  Same shape as pairing/bls12381/kilic.
package b200

/*
#cgo LDFLAGS: -lb2kyber
#include <stdlib.h>
#include <b2kyber.h>
*/
import "C"

import (
	"errors"
	"runtime"
	"sync"
	"unsafe"
)

// engine wraps one b2k context (one CUDA stream + scratch).  Calls on one context are serialised.
type engine struct {
	mu  sync.Mutex
	ctx *C.b2k_ctx
}

var (
	engOnce sync.Once
	eng     *engine
	engErr  error
)

// getEngine lazily creates the process-wide engine on device 0 (kyber has no context object; the
// adapters construct their third-party engines on every call, kilic/g1.go:50).
func getEngine() *engine {
	engOnce.Do(func() {
		var c *C.b2k_ctx
		if rc := C.b2k_create(0, &c); rc != 0 {
			engErr = errors.New("b200: no sm_100 device available (there is no CPU fallback)")
			return
		}
		eng = &engine{ctx: c}
		runtime.SetFinalizer(eng, func(e *engine) { C.b2k_destroy(e.ctx) })
	})
	if engErr != nil {
		panic(engErr)
	}
	return eng
}

func (e *engine) lastError() string { return C.GoString(C.b2k_last_error(e.ctx)) }

func ptr(b []byte) *C.uint8_t { return (*C.uint8_t)(unsafe.Pointer(&b[0])) }

// check panics on engine errors: kyber's Point methods have no error returns and panic on misuse.
func (e *engine) check(rc C.int) {
	if rc != 0 {
		panic("b200: " + e.lastError())
	}
}

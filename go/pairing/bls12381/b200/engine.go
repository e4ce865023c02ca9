//go:build b200

// Package b200 is a BLS12-381 backend for kyber whose group arithmetic runs on NVIDIA B200 GPUs through
// libb2kyber.so (include/b2kyber.h).  Same shape as pairing/bls12381/kilic: Suite, G1Elt, G2Elt, GTElt, groupBls.
// There is no CPU fallback: without an sm_100 device the first operation panics.
package b200

/*
#cgo LDFLAGS: -lb2kyber
#include <stdlib.h>
#include <b2kyber.h>
*/
import "C"

import (
	"errors"
	"os"
	"runtime"
	"strconv"
	"sync"
	"unsafe"
)

// engine wraps one b2k context (one CUDA stream + scratch arena).  A context serves one call at a time;
// the pool below hands every goroutine its own, so concurrent use of a suite (kyber's contract, see
// pairing/bls12381/bls12381_test.go:476-497 TestRacePairings) runs on as many streams as there are callers.
type engine struct {
	ctx *C.b2k_ctx
}

var (
	poolOnce sync.Once
	pool     chan *engine
	poolErr  error
)

// poolSize: B2K_CONTEXTS contexts (default: min(GOMAXPROCS, 8)) on device B2K_DEVICE (default 0).
func poolSize() int {
	if v, err := strconv.Atoi(os.Getenv("B2K_CONTEXTS")); err == nil && v > 0 {
		return v
	}
	n := runtime.GOMAXPROCS(0)
	if n > 8 {
		n = 8
	}
	return n
}

func initPool() {
	dev, _ := strconv.Atoi(os.Getenv("B2K_DEVICE"))
	n := poolSize()
	pool = make(chan *engine, n)
	for i := 0; i < n; i++ {
		var c *C.b2k_ctx
		if rc := C.b2k_create(C.int(dev), &c); rc != 0 {
			poolErr = errors.New("b200: no sm_100 device available (there is no CPU fallback)")
			return
		}
		pool <- &engine{ctx: c}
	}
}

// acquire blocks until a context is free; release returns it.  (Contexts live for the life of the process.)
func acquire() *engine {
	poolOnce.Do(initPool)
	if poolErr != nil {
		panic(poolErr)
	}
	return <-pool
}

func (e *engine) release() { pool <- e }

func (e *engine) lastError() string { return C.GoString(C.b2k_last_error(e.ctx)) }

func ptr(b []byte) *C.uint8_t { return (*C.uint8_t)(unsafe.Pointer(&b[0])) }

// check panics on engine errors: kyber's Point methods have no error returns and panic on misuse.
func (e *engine) check(rc C.int) {
	if rc != 0 {
		msg := e.lastError()
		e.release()
		panic("b200: " + msg)
	}
}

// with runs f on a pooled context.
func with(f func(e *engine)) {
	e := acquire()
	f(e)
	e.release()
}

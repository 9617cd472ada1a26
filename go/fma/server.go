package fma

// HTTP surface B1 (SURVEY.md §8b) served straight from Go for an engine owner that is not vLLM: the three routes the
// dual-pods controller calls on the instance port (pkg/controller/dual-pods/inference-server.go:1329-1339 POST /sleep,
// :1118-1137 POST /wake_up, :1595-1607 GET /is_sleeping; executable spec cmd/test-server/main.go:69-91).
//
// NOT COMPILED OR TESTED HERE (no Go toolchain in the build image); the Python mirror with the same semantics
// (llm-d-fast-model-actuation_b200/server.py) is the one under test (tests/test_server_contract.py).

import (
	"encoding/json"
	"net/http"
	"strconv"
	"sync"
)

// Server owns one engine per rank of an inference server and fans /sleep and /wake_up out to all of them, answering
// 200 only after every rank has finished (vllm:v1/executor/abstract.py:322-360 waits for all workers).
type Server struct {
	mu      sync.Mutex // one actuation at a time, like the engine-core RPC queue
	Engines []*Engine
	Weights uint64 // tag mask of "weights"
	Tier    int
}

func (s *Server) each(f func(*Engine) error) error {
	errs := make([]error, len(s.Engines))
	var wg sync.WaitGroup
	for i, e := range s.Engines {
		wg.Add(1)
		go func(i int, e *Engine) { // ranks are independent: no collective on this path
			defer wg.Done()
			errs[i] = f(e)
		}(i, e)
	}
	wg.Wait()
	for _, err := range errs {
		if err != nil {
			return err
		}
	}
	return nil
}

// Handler returns the mux with the three dev-mode routes.
func (s *Server) Handler() http.Handler {
	mux := http.NewServeMux()
	mux.HandleFunc("/sleep", func(w http.ResponseWriter, r *http.Request) {
		if r.Method != http.MethodPost {
			http.Error(w, "method not allowed", http.StatusMethodNotAllowed)
			return
		}
		level := 1 // the controller never sends level; vLLM's default is 1 (api_router.py:22-33)
		if v := r.URL.Query().Get("level"); v != "" {
			n, err := strconv.Atoi(v)
			if err != nil {
				http.Error(w, "bad level", http.StatusUnprocessableEntity)
				return
			}
			level = n
		}
		mask := s.Weights // level 1: offload ("weights",); level 2: offload nothing (gpu_worker.py:169-170)
		if level != 1 {
			mask = 0
		}
		s.mu.Lock()
		err := s.each(func(e *Engine) error { return e.Sleep(mask, s.Tier) }) // sleeping twice is a no-op inside the engine
		s.mu.Unlock()
		if err != nil {
			http.Error(w, err.Error(), http.StatusInternalServerError)
			return
		}
		w.WriteHeader(http.StatusOK) // exactly 200, empty body (inference-server.go:1335)
	})
	mux.HandleFunc("/wake_up", func(w http.ResponseWriter, r *http.Request) {
		if r.Method != http.MethodPost {
			http.Error(w, "method not allowed", http.StatusMethodNotAllowed)
			return
		}
		tags := r.URL.Query()["tags"] // repeated ?tags=weights&tags=kv_cache; none = wake everything
		s.mu.Lock()
		err := s.each(func(e *Engine) error {
			var mask uint64
			for _, t := range tags {
				id, terr := e.Tag(t)
				if terr != nil {
					return terr
				}
				mask |= 1 << uint(id)
			}
			return e.WakeUp(mask) // idempotent: the controller retries on its 5 s timeout (inference-server.go:1699-1716)
		})
		s.mu.Unlock()
		if err != nil {
			http.Error(w, err.Error(), http.StatusInternalServerError)
			return
		}
		w.WriteHeader(http.StatusOK)
	})
	mux.HandleFunc("/is_sleeping", func(w http.ResponseWriter, r *http.Request) {
		sleeping := false
		for _, e := range s.Engines {
			sleeping = sleeping || e.IsSleeping()
		}
		w.Header().Set("Content-Type", "application/json")
		_ = json.NewEncoder(w).Encode(map[string]bool{"is_sleeping": sleeping}) // api.SleepState, pkg/api/interface.go:131-133
	})
	return mux
}

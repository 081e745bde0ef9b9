# Convenience targets (the driver uses __graft_entry__.build(); these call the same recipes).
PY ?= python

build:            ## libmgb200.so (nvcc, sm_100a), CPU oracle, oracle/_ref when /root/reference is mounted
	$(PY) -c "import __graft_entry__ as g; g.build()"

test-cpu:         ## everything that runs without a GPU (oracle pins, C-ABI exports, emulation of tile / cluster-tail kernels, slab schedules over gloo)
	$(PY) -m pytest tests -x -q -m "not gpu"

test-gpu:         ## parity tests proper (needs a B200)
	$(PY) -m pytest tests -x -q -m gpu

test-gpu-optin:   ## parity tests of the opt-in paths (tile kernels, zero-guess chain, cluster tail, comm-avoiding slabs)
	MGB200_TEST_OPTIN=1 $(PY) -m pytest tests/test_optin_gpu.py -x -q -m gpu

example: build    ## the reference's main() over the C ABI
	g++ -O2 -std=c++17 -Iinclude examples/poisson_main.cpp -o poisson_main \
	    -Lmultigrid_nikhil_c-_b200/lib -lmgb200 -Wl,-rpath,$(CURDIR)/multigrid_nikhil_c-_b200/lib

bench:
	$(PY) bench.py

.PHONY: build test-cpu test-gpu test-gpu-optin example bench

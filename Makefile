# Build / run entry points with the shape of the reference's Makefile (`make`, `make run`, `make clean`;
# reference: Makefile:36-63 builds bin/solver<N>.out and `make run<N>` runs it on input/sample.txt).
# Same commands as simplex_method_gpu_b200/_build.py (python -c "import __graft_entry__ as g; g.build()").
NVCC      := nvcc
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVCCFLAGS := $(ARCH) -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -pthread
PKG       := simplex_method_gpu_b200
CSRC      := $(PKG)/csrc
LIB       := $(PKG)/libb200lp.so
CLI       := bin/solver.out
INPUT     := tests/golden/sample.txt

all: $(LIB) $(CLI)

$(LIB): $(CSRC)/engine.cu $(CSRC)/kernels.cuh $(CSRC)/lp_io.cpp include/b200lp.h include/b200lp_io.h
	$(NVCC) $(NVCCFLAGS) -shared -o $@ $(CSRC)/engine.cu $(CSRC)/lp_io.cpp

$(CLI): $(CSRC)/solver_main.cpp include/b200lp.h include/b200lp_io.h $(LIB)
	@mkdir -p bin
	$(NVCC) -O2 -std=c++17 -o $@ $(CSRC)/solver_main.cpp -I include -L $(PKG) -lb200lp \
		-Xlinker -rpath -Xlinker '$$ORIGIN/../$(PKG)'

# `make run` = the reference's `make run4` on the same LP (needs a B200)
run: $(CLI)
	@echo "--- Running: $(CLI) $(INPUT) ---"
	@./$(CLI) $(INPUT)

oracle:
	$(MAKE) -C oracle
	bash oracle/make_ref.sh

test:
	python -m pytest tests -q -m "not gpu"

test-gpu:
	python -m pytest tests -q -m gpu

bench:
	python bench.py

clean:
	rm -rf bin $(LIB) oracle/liboracle.so

.PHONY: all run oracle test test-gpu bench clean

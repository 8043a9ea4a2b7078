# Builds libmrq.so (the C-ABI engine, sm_100a only) in-tree and the CPU oracle (test infrastructure).
NVCC ?= /usr/local/cuda/bin/nvcc
ARCH := -gencode arch=compute_100a,code=sm_100a
NVFLAGS := -O3 -std=c++17 -lineinfo $(ARCH) -Xcompiler -fPIC,-Wall,-Wextra -Xptxas -v
LIB := raftsql_b200/libmrq.so
SRC := raftsql_b200/csrc/mrq_engine.cu
HDR := raftsql_b200/csrc/mrq_kernels.cuh include/mrq.h include/mrq_trace.h

all: $(LIB) oracle

$(LIB): $(SRC) $(HDR)
	$(NVCC) $(NVFLAGS) -shared -o $@ $(SRC) -ldl 2> build_ptxas.log || (cat build_ptxas.log; exit 1)

oracle:
	$(MAKE) -C oracle liboracle.so

clean:
	rm -f $(LIB) build_ptxas.log
	$(MAKE) -C oracle clean
.PHONY: all oracle clean

# Builds, in-tree: libmrq.so (the C-ABI engine, sm_100a only), libraftpipe.so (the C++ host side above the
# C-ABI: channels, host node, WAL, NewRaftPipe), the CPU oracle (test infrastructure) and the C++ scenario test.
NVCC ?= /usr/local/cuda/bin/nvcc
CXX ?= g++
ARCH := -gencode arch=compute_100a,code=sm_100a
NVFLAGS := -O3 -std=c++17 -lineinfo $(ARCH) -Xcompiler -fPIC,-Wall,-Wextra -Xptxas -v
CXXFLAGS := -O2 -std=c++17 -fPIC -Wall -Wextra -pthread
ROOT := $(abspath .)
LIB := raftsql_b200/libmrq.so
HOSTLIB := raftsql_b200/libraftpipe.so
SRC := raftsql_b200/csrc/mrq_engine.cu raftsql_b200/csrc/mrq_pack8_rows.cpp
HDR := raftsql_b200/csrc/mrq_kernels.cuh include/mrq.h include/mrq_trace.h include/mrq_packed8.h
HOSTSRC := raftsql_b200/csrc/host/hostnode.cpp raftsql_b200/csrc/host/raftpipe.cpp
HOSTHDR := raftsql_b200/csrc/host/chan.hpp raftsql_b200/csrc/host/hostnode.hpp raftsql_b200/csrc/host/raftpipe.hpp include/mrq.h
CPPTEST := tests/cpp/raftpipe_test

all: $(LIB) $(HOSTLIB) oracle $(CPPTEST)

$(LIB): $(SRC) $(HDR)
	$(NVCC) $(NVFLAGS) -shared -o $@ $(SRC) -ldl 2> build_ptxas.log || (cat build_ptxas.log; exit 1)

$(HOSTLIB): $(HOSTSRC) $(HOSTHDR) $(LIB)
	$(CXX) $(CXXFLAGS) -shared -o $@ $(HOSTSRC) -L$(ROOT)/raftsql_b200 -lmrq -Wl,-rpath,'$$ORIGIN'

oracle:
	$(MAKE) -C oracle liboracle.so

$(CPPTEST): tests/cpp/raftpipe_test.cpp $(HOSTLIB) oracle
	$(CXX) $(CXXFLAGS) -o $@ tests/cpp/raftpipe_test.cpp -L$(ROOT)/raftsql_b200 -lraftpipe -lmrq -L$(ROOT)/oracle -loracle \
	  -Wl,-rpath,$(ROOT)/raftsql_b200 -Wl,-rpath,$(ROOT)/oracle

# The C++ host side and its scenario test under the sanitizers (CPU suite; oracle core only — no device code runs).
# SAN_CXX: the first of $(CXX), g++, /usr/bin/g++ that ships the sanitizer runtimes (`make san_cxx` prints it).
SAN_CXX := $(shell for c in $(CXX) g++ /usr/bin/g++; do f=`$$c -print-file-name=libtsan.so 2>/dev/null`; \
	if [ -n "$$f" ] && [ "$$f" != libtsan.so ]; then echo $$c; break; fi; done)
SAN_SRC := tests/cpp/raftpipe_test.cpp $(HOSTSRC) oracle/raft_oracle.c
san_cxx:
	@echo $(SAN_CXX)
tests/cpp/raftpipe_test_asan: $(SAN_SRC) $(HOSTHDR) $(LIB)
	$(SAN_CXX) -O1 -g -std=c++17 -pthread -fsanitize=address,undefined -fno-sanitize-recover=undefined -o $@ $(SAN_SRC) \
	  -L$(ROOT)/raftsql_b200 -lmrq -Wl,-rpath,$(ROOT)/raftsql_b200
tests/cpp/raftpipe_test_tsan: $(SAN_SRC) $(HOSTHDR) $(LIB)
	$(SAN_CXX) -O1 -g -std=c++17 -pthread -fsanitize=thread -o $@ $(SAN_SRC) \
	  -L$(ROOT)/raftsql_b200 -lmrq -Wl,-rpath,$(ROOT)/raftsql_b200

# The per-group tick functions of mrq_kernels.cuh compiled for the HOST and run against the oracle (CPU suite):
# the device arithmetic, same source, checked where there is no GPU.  The _san build adds ASan + UBSan.
CUDA_INC ?= /usr/local/cuda/include
TICKHOST_SRC := tests/cpp/tick_host_test.cpp oracle/raft_oracle.c
TICKHOST_DEP := $(TICKHOST_SRC) tests/cpp/device_on_host.hpp oracle/raft_oracle.h $(HDR)
tests/cpp/tick_host_test: $(TICKHOST_DEP)
	$(CXX) -std=c++17 -O1 -I$(CUDA_INC) -pthread -o $@ $(TICKHOST_SRC)
tests/cpp/tick_host_test_san: $(TICKHOST_DEP)
	$(SAN_CXX) -std=c++17 -O0 -fsanitize=address,undefined -fno-sanitize-recover=undefined -I$(CUDA_INC) -pthread -o $@ $(TICKHOST_SRC)

clean:
	rm -f $(LIB) $(HOSTLIB) $(CPPTEST) tests/cpp/raftpipe_test_asan tests/cpp/raftpipe_test_tsan build_ptxas.log \
	  tests/cpp/tick_host_test tests/cpp/tick_host_test_san
	$(MAKE) -C oracle clean
.PHONY: all oracle clean san_cxx

# Builds libqpswift_b200.so (host C++ + sm_100a CUDA) in-tree, plus the test oracles.
#   make            -> apf_quadruped_b200/lib/libqpswift_b200.so (incl. the pattern-specialised kernels generated into
#                      build/gen by apf_quadruped_b200/codegen.py), oracle port, oracle/_ref (if reference present)
# -fmad=false is part of the numerical contract (bit-exact with the reference's x86-64 arithmetic).
NVCC     ?= /usr/local/cuda/bin/nvcc
CXX      ?= g++
ARCH     := -gencode arch=compute_100a,code=sm_100a
NVFLAGS  := -O3 -std=c++17 $(ARCH) -lineinfo -fmad=false -Xcompiler -fPIC -Iinclude -Iapf_quadruped_b200/csrc
CXXFLAGS := -O2 -std=c++17 -fPIC -Iinclude -Iapf_quadruped_b200/csrc -I/usr/local/cuda/include
SRC      := apf_quadruped_b200/csrc
OBJ      := build/obj
LIB      := apf_quadruped_b200/lib/libqpswift_b200.so

HOST_SRCS := $(SRC)/amd_order.cpp $(SRC)/symbolic.cpp $(SRC)/program.cpp $(SRC)/stream_program.cpp $(SRC)/capi_plan.cpp $(SRC)/capi_qp.cpp $(SRC)/multi.cpp
HOST_OBJS := $(patsubst $(SRC)/%.cpp,$(OBJ)/%.o,$(HOST_SRCS))
CU_SRCS   := $(SRC)/batch.cu
CU_OBJS   := $(patsubst $(SRC)/%.cu,$(OBJ)/%.cu.o,$(CU_SRCS))
# pattern-specialised throughput kernels: `make gen` (needs the library itself, so it is a second
# pass: make && make gen && make) writes build/gen/stream_gen_<kind>.cu with apf_quadruped_b200/codegen.py
GEN       := build/gen
GEN_KINDS ?= trot stance crawl
GEN_SRCS  := $(wildcard $(GEN)/*.cu)
GEN_OBJS  := $(patsubst $(GEN)/%.cu,$(OBJ)/gen_%.cu.o,$(GEN_SRCS))

# Two passes: the generator needs the library's host side (plan + micro-programs), and the generated
# translation units are then compiled into the same library.
all:
	$(MAKE) $(LIB)
	$(MAKE) gen
	$(MAKE) $(LIB)
	$(MAKE) oracle
	$(MAKE) ctest

$(OBJ)/%.o: $(SRC)/%.cpp $(wildcard $(SRC)/*.h) $(wildcard include/*.h include/qpSWIFT/*.h)
	@mkdir -p $(OBJ)
	$(CXX) $(CXXFLAGS) -c $< -o $@

$(OBJ)/%.cu.o: $(SRC)/%.cu $(wildcard $(SRC)/*.h $(SRC)/*.cuh) $(wildcard include/*.h include/qpSWIFT/*.h)
	@mkdir -p $(OBJ)
	$(NVCC) $(NVFLAGS) -Xptxas -v -c $< -o $@ 2> $(OBJ)/$*.ptxas.log || (cat $(OBJ)/$*.ptxas.log; false)

$(OBJ)/gen_%.cu.o: $(GEN)/%.cu $(wildcard $(SRC)/*.h $(SRC)/*.cuh) $(wildcard include/*.h include/qpSWIFT/*.h)
	@mkdir -p $(OBJ)
	$(NVCC) $(NVFLAGS) -Xptxas -v -c $< -o $@ 2> $(OBJ)/gen_$*.ptxas.log || (tail -50 $(OBJ)/gen_$*.ptxas.log; false)

gen:
	python -m apf_quadruped_b200.codegen --out $(GEN) $(GEN_KINDS)
	python -m apf_quadruped_b200.codegen_warp --out $(GEN) $(GEN_KINDS)

$(LIB): $(HOST_OBJS) $(CU_OBJS) $(GEN_OBJS)
	@mkdir -p $(dir $(LIB))
	$(NVCC) $(ARCH) -shared -o $@ $^ -cudart shared -ldl -lpthread

oracle:
	$(MAKE) -C oracle all

# the C caller of the multi-GPU entry points (tests/test_multi_gpu.py runs it on the GPU box)
ctest: tests/c/multi_gpu_test
tests/c/multi_gpu_test: tests/c/multi_gpu_test.c $(LIB) include/qpswift_b200.h
	gcc -O2 -std=gnu99 -Wall -Iinclude -o $@ $< -L$(dir $(LIB)) -lqpswift_b200 -Wl,-rpath,'$$ORIGIN/../../apf_quadruped_b200/lib'

clean:
	rm -rf build $(LIB) tests/c/multi_gpu_test
	$(MAKE) -C oracle clean

.PHONY: all oracle clean gen ctest

# Builds libqpswift_b200.so (host C++ + sm_100a CUDA) in-tree, plus the test oracles.
#   make            -> apf_quadruped_b200/lib/libqpswift_b200.so, oracle port, oracle/_ref (if reference present)
# -fmad=false is part of the numerical contract (bit-exact with the reference's x86-64 arithmetic).
NVCC     ?= /usr/local/cuda/bin/nvcc
CXX      ?= g++
ARCH     := -gencode arch=compute_100a,code=sm_100a
NVFLAGS  := -O3 -std=c++17 $(ARCH) -lineinfo -fmad=false -Xcompiler -fPIC -Iinclude -Iapf_quadruped_b200/csrc
CXXFLAGS := -O2 -std=c++17 -fPIC -Iinclude -Iapf_quadruped_b200/csrc -I/usr/local/cuda/include
SRC      := apf_quadruped_b200/csrc
OBJ      := build/obj
LIB      := apf_quadruped_b200/lib/libqpswift_b200.so

HOST_SRCS := $(SRC)/amd_order.cpp $(SRC)/symbolic.cpp $(SRC)/program.cpp $(SRC)/stream_program.cpp $(SRC)/capi_plan.cpp $(SRC)/capi_qp.cpp
HOST_OBJS := $(patsubst $(SRC)/%.cpp,$(OBJ)/%.o,$(HOST_SRCS))
CU_SRCS   := $(SRC)/batch.cu
CU_OBJS   := $(patsubst $(SRC)/%.cu,$(OBJ)/%.cu.o,$(CU_SRCS))

all: $(LIB) oracle

$(OBJ)/%.o: $(SRC)/%.cpp $(wildcard $(SRC)/*.h) $(wildcard include/*.h include/qpSWIFT/*.h)
	@mkdir -p $(OBJ)
	$(CXX) $(CXXFLAGS) -c $< -o $@

$(OBJ)/%.cu.o: $(SRC)/%.cu $(wildcard $(SRC)/*.h $(SRC)/*.cuh) $(wildcard include/*.h include/qpSWIFT/*.h)
	@mkdir -p $(OBJ)
	$(NVCC) $(NVFLAGS) -Xptxas -v -c $< -o $@ 2> $(OBJ)/$*.ptxas.log || (cat $(OBJ)/$*.ptxas.log; false)

$(LIB): $(HOST_OBJS) $(CU_OBJS)
	@mkdir -p $(dir $(LIB))
	$(NVCC) $(ARCH) -shared -o $@ $^ -cudart shared

oracle:
	$(MAKE) -C oracle all

clean:
	rm -rf build $(LIB)
	$(MAKE) -C oracle clean

.PHONY: all oracle clean

# Builds libfoodrec_b200.so (CUDA kernels + C ABI, sm_100a only) in-tree.
PKG   := multi-modal-food-recommendation_b200
CSRC  := $(PKG)/csrc
NVCC  ?= /usr/local/cuda/bin/nvcc
ARCH  := -gencode arch=compute_100a,code=sm_100a
NVFLAGS := -O3 -std=c++17 -lineinfo $(ARCH) -Xcompiler -fPIC -Iinclude -I$(CSRC) --expt-relaxed-constexpr $(EXTRA)
SRCS  := $(wildcard $(CSRC)/*.cu)
OBJS  := $(patsubst $(CSRC)/%.cu,build/%.o,$(SRCS))
LIB   := $(PKG)/libfoodrec_b200.so

all: $(LIB)

build/%.o: $(CSRC)/%.cu $(wildcard $(CSRC)/*.cuh) include/foodrec_b200.h
	@mkdir -p build
	$(NVCC) $(NVFLAGS) -Xptxas -v -c $< -o $@ 2> build/$*.ptxas.log || (cat build/$*.ptxas.log; false)

$(LIB): $(OBJS)
	$(NVCC) -shared $(ARCH) -o $@ $(OBJS) -lcudart -ldl

clean:
	rm -rf build $(LIB)
.PHONY: all clean

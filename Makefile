# Builds the C-ABI library (the product) and the standalone GPU self-test.
NVCC ?= nvcc
ARCH := -gencode arch=compute_100a,code=sm_100a
NVFLAGS := $(ARCH) -lineinfo -O3 -std=c++17
CSRC := mrclip_b200/csrc
HDRS := $(CSRC)/ptx.cuh $(CSRC)/peer_sync.cuh $(CSRC)/peer_kernels.cuh $(CSRC)/tile_kernel.cuh $(CSRC)/gemm_kernel.cuh $(CSRC)/gemm2_kernel.cuh $(CSRC)/aux_kernels.cuh include/mrclip.h

all: mrclip_b200/libmrclip.so mrclip_b200/selftest

mrclip_b200/libmrclip.so: $(CSRC)/mrclip_cabi.cu $(HDRS)
	$(NVCC) $(NVFLAGS) --shared -Xcompiler -fPIC -o $@ $<

mrclip_b200/selftest: $(CSRC)/selftest.cu mrclip_b200/libmrclip.so include/mrclip.h
	$(NVCC) $(NVFLAGS) -Xcompiler -fopenmp -o $@ $< -Lmrclip_b200 -lmrclip -Xlinker -rpath='$$ORIGIN'

clean:
	rm -f mrclip_b200/libmrclip.so mrclip_b200/selftest

.PHONY: all clean

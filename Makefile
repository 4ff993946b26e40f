# Builds libopz.so (the C-ABI shared library) for sm_100a, and the C oracle.
NVCC ?= nvcc
PKG := oceanparameterizations.jl_b200
CSRC := $(PKG)/csrc
NVFLAGS := -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-Wall -Xptxas -v
SRCS := $(CSRC)/opz_capi.cu $(CSRC)/opz_fp32.cu $(CSRC)/opz_bptt.cu $(CSRC)/opz_tc.cu $(CSRC)/opz_tc2.cu $(CSRC)/opz_tc_probe.cu
OBJS := $(SRCS:.cu=.o)
HDRS := $(wildcard $(CSRC)/*.cuh) $(wildcard $(CSRC)/*.h) include/opz.h

all: $(PKG)/libopz.so oracle/_build/liboracle.so

$(CSRC)/%.o: $(CSRC)/%.cu $(HDRS)
	$(NVCC) $(NVFLAGS) -c $< -o $@ 2> $@.ptxas.log || (cat $@.ptxas.log; exit 1)

$(PKG)/libopz.so: $(OBJS)
	$(NVCC) -Wno-deprecated-gpu-targets -shared -o $@ $(OBJS) -lcudart_static -lpthread -ldl -lrt

oracle/_build/liboracle.so: oracle/nde_oracle.c
	mkdir -p oracle/_build
	gcc -O3 -march=x86-64-v3 -fopenmp -fPIC -shared -o $@ $< -lm

clean:
	rm -f $(CSRC)/*.o $(CSRC)/*.log $(PKG)/libopz.so oracle/_build/liboracle.so
.PHONY: all clean

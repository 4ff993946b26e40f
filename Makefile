# Builds libopz.so (the C-ABI shared library) for sm_100a, the probe library used only by tests, and the C oracle.
# Objects and ptxas logs go to build/ (git-ignored); the .so files stay in-tree so that they travel to the GPU box.
NVCC ?= nvcc
PKG := oceanparameterizations.jl_b200
CSRC := $(PKG)/csrc
BUILD := build
NVFLAGS := -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-Wall -Xptxas -v
NAMES := opz_capi opz_comm opz_fp32 opz_bptt opz_tc2 opz_gemm
OBJS := $(NAMES:%=$(BUILD)/%.o)
HDRS := $(wildcard $(CSRC)/*.cuh) $(wildcard $(CSRC)/*.h) include/opz.h

all: $(PKG)/libopz.so tests/libopz_probe.so oracle/_build/liboracle.so

$(BUILD)/%.o: $(CSRC)/%.cu $(HDRS)
	@mkdir -p $(BUILD)
	$(NVCC) $(NVFLAGS) -c $< -o $@ 2> $(BUILD)/$*.ptxas.log || (cat $(BUILD)/$*.ptxas.log; exit 1)

$(PKG)/libopz.so: $(OBJS)
	$(NVCC) -Wno-deprecated-gpu-targets -shared -o $@ $(OBJS) -lcudart_static -lpthread -ldl -lrt

# tcgen05 / TMEM micro-probes (operand-convention and issue-rate checks): test infrastructure, not shipped in libopz.so
tests/libopz_probe.so: $(BUILD)/opz_tc_probe.o $(PKG)/libopz.so
	$(NVCC) -Wno-deprecated-gpu-targets -shared -o $@ $(BUILD)/opz_tc_probe.o -L$(PKG) -lopz -Xlinker -rpath,'$$ORIGIN/../$(PKG)' -lcudart_static -lpthread -ldl -lrt

$(BUILD)/opz_tc_probe.o: tests/csrc/opz_tc_probe.cu $(HDRS)
	@mkdir -p $(BUILD)
	$(NVCC) $(NVFLAGS) -I$(CSRC) -c $< -o $@ 2> $(BUILD)/opz_tc_probe.ptxas.log || (cat $(BUILD)/opz_tc_probe.ptxas.log; exit 1)

oracle/_build/liboracle.so: oracle/nde_oracle.c
	mkdir -p oracle/_build
	gcc -O3 -march=x86-64-v3 -fopenmp -fPIC -shared -o $@ $< -lm

clean:
	rm -rf $(BUILD) $(PKG)/libopz.so tests/libopz_probe.so oracle/_build/liboracle.so
.PHONY: all clean

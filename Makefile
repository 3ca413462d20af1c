# Builds the C-ABI library in-tree (the .so travels to the GPU box with the gpurun snapshot).
NVCC      ?= nvcc
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVCCFLAGS := -O3 -Xptxas -v -std=c++17 -lineinfo $(ARCH) -Xcompiler -fPIC -Xcompiler -Wall -Xcompiler -Wno-unknown-pragmas --expt-relaxed-constexpr
CSRC      := lfm_quant_b200/csrc
SRCS      := $(CSRC)/lfmq_api.cu $(CSRC)/kernels_simt.cu $(CSRC)/lstm_tc.cu $(CSRC)/rnn_tc.cu
OBJS      := $(SRCS:.cu=.o)
LIB       := lfm_quant_b200/_lfmq.so

all: $(LIB)

$(CSRC)/%.o: $(CSRC)/%.cu $(wildcard $(CSRC)/*.h $(CSRC)/*.cuh) include/lfmq.h
	$(NVCC) $(NVCCFLAGS) -c $< -o $@

$(LIB): $(OBJS)
	$(NVCC) $(ARCH) -shared -o $@ $(OBJS) -lcudart

clean:
	rm -f $(OBJS) $(LIB)

tools: tools/tc_probe
tools/tc_probe: tools/tc_probe.cu $(CSRC)/sm100.cuh
	$(NVCC) -O3 -std=c++17 -lineinfo $(ARCH) -o $@ $<

# A/B builds of kernel variants: `make alt ALT_FLAGS="-DLFMQ_BWD_LATE_C0=1"` -> lfm_quant_b200/_lfmq_alt.so (LFMQ_LIB_PATH selects it)
alt:
	mkdir -p build/alt
	for f in lfmq_api kernels_simt lstm_tc rnn_tc; do $(NVCC) $(NVCCFLAGS) $(ALT_FLAGS) -c $(CSRC)/$$f.cu -o build/alt/$$f.o || exit 1; done
	$(NVCC) $(ARCH) -shared -o lfm_quant_b200/_lfmq_alt.so build/alt/lfmq_api.o build/alt/kernels_simt.o build/alt/lstm_tc.o build/alt/rnn_tc.o -lcudart

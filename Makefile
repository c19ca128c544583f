# Build libacino_b200.so (sm_100a only) and the CPU oracle's C restatement.
NVCC      ?= /usr/local/cuda/bin/nvcc
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVCCFLAGS := -O3 -std=c++17 $(ARCH) -lineinfo -Xcompiler -fPIC -cudart static --expt-relaxed-constexpr
CSRC      := acinoset_b200/csrc
SRCS      := $(wildcard $(CSRC)/*.cu)
HDRS      := $(wildcard $(CSRC)/*.cuh) include/acino_b200.h
LIB       := acinoset_b200/libacino_b200.so

all: $(LIB) oracle

$(LIB): $(SRCS) $(HDRS)
	$(NVCC) $(NVCCFLAGS) -shared -o $@ $(SRCS)

ptxas-info: $(SRCS) $(HDRS)
	$(NVCC) $(NVCCFLAGS) -Xptxas -v -shared -o /tmp/libacino_b200_ptxas.so $(SRCS)

# A/B build for scripts/ab_env.sh / ab_e2e.sh: the ACINO_FTE_VARIANT / ACINO_FTE_CTAS /
# ACINO_FTE_SMEM_PAD / ACINO_E2E_CHUNK environment switches exist only in this build (the product library has none).
experiments: $(SRCS) $(HDRS)
	mkdir -p scratch
	$(NVCC) $(NVCCFLAGS) -DACINO_EXPERIMENTS -DACINO_MAXC_STAGE=6 -shared -o scratch/libacino_b200_experiments.so $(SRCS)

oracle:
	$(MAKE) -C oracle

clean:
	rm -f $(LIB)
	$(MAKE) -C oracle clean

.PHONY: all oracle clean ptxas-info experiments

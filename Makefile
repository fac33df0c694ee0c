# Builds libphare_b200.so (hand-written sm_100a kernels behind the C ABI of include/phare_b200.h)
# and the test-only CPU oracle.  nvcc cross-compiles without a GPU.
NVCC      ?= /usr/local/cuda/bin/nvcc
ARCH      := -gencode arch=compute_100a,code=sm_100a
# -fmad=false: products and sums are rounded separately, like the reference's default CPU build;
# the few deliberate fusions are explicit fma() calls.
NVCCFLAGS := $(ARCH) -O3 -std=c++17 -lineinfo -fmad=false -Xcompiler -fPIC -Xptxas -v --expt-relaxed-constexpr $(EXTRA)
SRC       := $(wildcard phare_b200/csrc/*.cu)
# tuning builds: make lib BUILD=build_v1 LIB=phare_b200/lib/libphare_b200_v1.so EXTRA=-DPHB_TILE_BS=256  (load with PHB_LIB=...)
BUILD     ?= build
OBJ       := $(patsubst phare_b200/csrc/%.cu,$(BUILD)/%.o,$(SRC))
LIB       ?= phare_b200/lib/libphare_b200.so

HOSTLIB   := phare_b200/lib/libphare_b200_host.so

all: lib host oracle

lib: $(LIB)

# the C++ level driver (include/phare_b200/solver_ppc.hpp) behind a small C ABI: plain g++, CUDA only through $(LIB)
host: $(HOSTLIB)
$(HOSTLIB): phare_b200/host/host_api.cpp $(wildcard include/phare_b200/*.hpp) include/phare_b200.h $(LIB)
	g++ -std=c++20 -O2 -fPIC -shared -Iinclude -o $@ phare_b200/host/host_api.cpp -Lphare_b200/lib -lphare_b200 \
	    -Wl,-rpath,'$$ORIGIN' -Wl,-rpath,/usr/local/cuda/lib64

$(BUILD)/%.o: phare_b200/csrc/%.cu $(wildcard phare_b200/csrc/*.cuh) include/phare_b200.h
	@mkdir -p $(BUILD)
	$(NVCC) $(NVCCFLAGS) -c $< -o $@ 2> $(BUILD)/$*.ptxas.log || (cat $(BUILD)/$*.ptxas.log; false)

$(LIB): $(OBJ)
	@mkdir -p phare_b200/lib
	$(NVCC) $(ARCH) -shared -o $@ $(OBJ) -lcudart

oracle:
	$(MAKE) -C oracle all

clean:
	rm -rf $(BUILD) $(LIB)
	$(MAKE) -C oracle clean
.PHONY: all lib host oracle clean

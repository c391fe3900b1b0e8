# Builds everything in-tree (no install step; the .so files travel with the repo snapshot to the GPU box):
#   pbrlab_b200/lib/libpbrgpu.so        the C ABI (include/pbrgpu.h): CUDA kernels for sm_100a + host-side BVH build
#   pbrlab_b200/lib/libpbrlab_host.so   C++ mirror of the reference API (Scene / Render / loaders) + C shim for ctypes
#   pbrlab_b200/lib/pbrlab-cli          command-line twin of the reference's pbrlab-cli, with --width/--height/--spp
#   tests/host_emul/libpbr_emul.so      TEST ONLY: the device headers compiled by g++ (no GPU in the dev container)
#   oracle/...                          TEST ONLY: see oracle/Makefile
NVCC     ?= /usr/local/cuda/bin/nvcc
HOSTCXX  := /usr/bin/g++
ARCH     := -gencode arch=compute_100a,code=sm_100a
# -fmad=false: the reference is compiled without FMA contraction; fused ops are spelled explicitly (device/common.cuh)
NVFLAGS  := $(ARCH) -O3 -lineinfo -fmad=false -std=c++17 -ccbin $(HOSTCXX) -Xcompiler -fPIC,-Wall,-Wno-unused-function \
            -Xptxas -v --expt-relaxed-constexpr
CXXFLAGS := -O2 -std=c++17 -fPIC -Wall -ffp-contract=off
LIB      := pbrlab_b200/lib
CSRC     := pbrlab_b200/csrc
HOST     := pbrlab_b200/host
DEVHDRS  := $(wildcard $(CSRC)/device/*.cuh) $(CSRC)/kat.cuh $(CSRC)/wavefront.cuh $(CSRC)/scene_host.h \
            $(CSRC)/bvh_builder.h $(CSRC)/job_split.h $(CSRC)/nccl_shim.h include/pbrgpu.h
HOSTSRCS := $(HOST)/scene.cc $(HOST)/render.cc $(HOST)/light-manager.cc $(HOST)/mesh/triangle-mesh.cc \
            $(HOST)/curve-util.cc $(HOST)/io/triangle-mesh-io.cc $(HOST)/io/image-io.cc $(HOST)/io/cyhair.cc $(HOST)/io/curve-mesh-io.cc \
            $(HOST)/pc-common.cc $(HOST)/c_api.cc
HOSTHDRS := $(wildcard $(HOST)/*.h $(HOST)/*/*.h)

.PHONY: all gpu host emul oracle clean
all: gpu host emul oracle

gpu: $(LIB)/libpbrgpu.so
host: $(LIB)/libpbrlab_host.so $(LIB)/pbrlab-cli
emul: tests/host_emul/libpbr_emul.so
oracle:
	$(MAKE) -C oracle all

$(LIB)/scene_host.o: $(CSRC)/scene_host.cc $(DEVHDRS)
	@mkdir -p $(LIB)
	$(HOSTCXX) $(CXXFLAGS) -c $< -o $@
$(LIB)/bvh_builder.o: $(CSRC)/bvh_builder.cc $(CSRC)/bvh_builder.h
	@mkdir -p $(LIB)
	$(HOSTCXX) $(CXXFLAGS) -c $< -o $@
$(LIB)/pbrgpu.o: $(CSRC)/pbrgpu.cu $(DEVHDRS)
	@mkdir -p $(LIB)
	$(NVCC) $(NVFLAGS) -c $< -o $@ 2> $(LIB)/ptxas.log || (cat $(LIB)/ptxas.log; false)
$(LIB)/libpbrgpu.so: $(LIB)/pbrgpu.o $(LIB)/scene_host.o $(LIB)/bvh_builder.o
	$(NVCC) $(ARCH) -ccbin $(HOSTCXX) -shared -o $@ $^ -lpthread

$(LIB)/libpbrlab_host.so: $(HOSTSRCS) $(HOSTHDRS) $(LIB)/libpbrgpu.so
	$(HOSTCXX) $(CXXFLAGS) -shared -o $@ $(HOSTSRCS) -L$(LIB) -lpbrgpu -Wl,-rpath,'$$ORIGIN' -lpthread -lz
$(LIB)/pbrlab-cli: $(HOST)/pbrlab-cli.cc $(LIB)/libpbrlab_host.so
	$(HOSTCXX) $(CXXFLAGS) -o $@ $(HOST)/pbrlab-cli.cc -L$(LIB) -lpbrlab_host -lpbrgpu -Wl,-rpath,'$$ORIGIN' -lpthread -lz

tests/host_emul/libpbr_emul.so: tests/host_emul/emul.cc $(CSRC)/scene_host.cc $(CSRC)/bvh_builder.cc $(DEVHDRS)
	$(HOSTCXX) $(CXXFLAGS) -Wno-unused-function -shared -o $@ tests/host_emul/emul.cc $(CSRC)/scene_host.cc \
	    $(CSRC)/bvh_builder.cc -lpthread

clean:
	rm -rf $(LIB) tests/host_emul/libpbr_emul.so
	$(MAKE) -C oracle clean

# Build of the B200 multi-view deconvolution library.
#
#   make            -> multiview-reconstruction_b200/libmvdecon.so   (product; CUDA, sm_100a only)
#   make hostemu    -> tests/host/libmvdecon_hostemu.so              (TEST ONLY: same kernel bodies run on the CPU)
#   make fft_emu_test
#
# MVD_LENGTHS="32 64 ..." restricts the instantiated FFT lengths (development builds).
NVCC      ?= /usr/local/cuda/bin/nvcc
CXX       ?= g++
PKG       := multiview-reconstruction_b200
CSRC      := $(PKG)/csrc
BUILD     := build
GEN       := $(BUILD)/gen
GENH      := $(BUILD)/hostemu/gen
NVFLAGS   := -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-fvisibility=hidden \
             --expt-relaxed-constexpr -Xptxas -v -diag-suppress 177,550
HOSTFLAGS := -std=c++17 -O1 -fPIC -fvisibility=hidden -DMVD_HOST_EMU -ffp-contract=off -I$(CSRC) -w
# lengths compiled into the host emulator (small ones for the tests + the c1..c3 tile lengths)
HOSTEMU_LENGTHS ?= 32 36 40 48 50 54 60 64 72 80 96 128 144 160
HDRS      := $(CSRC)/fft_codelets.cuh $(CSRC)/fft_passes.cuh $(CSRC)/backend.h $(CSRC)/len_ops_impl.cuh $(CSRC)/engine.h include/mvdecon.h

LENS      := $(shell MVD_LENGTHS="$(MVD_LENGTHS)" python3 $(CSRC)/gen_lengths.py $(GEN))
LENOBJ    := $(foreach n,$(LENS),$(BUILD)/obj/len_$(n).o)
OBJ       := $(LENOBJ) $(BUILD)/obj/registry.o $(BUILD)/obj/engine.o $(BUILD)/obj/pointwise.o $(BUILD)/obj/comm.o $(BUILD)/obj/psf_prep.o $(BUILD)/obj/tiff_io.o $(BUILD)/obj/n5_io.o $(BUILD)/obj/capi.o

HLENS     := $(shell MVD_LENGTHS="$(HOSTEMU_LENGTHS)" python3 $(CSRC)/gen_lengths.py $(GENH))
HLENOBJ   := $(foreach n,$(HLENS),$(BUILD)/hostemu/obj/len_$(n).o)
HOBJ      := $(HLENOBJ) $(BUILD)/hostemu/obj/registry.o $(BUILD)/hostemu/obj/engine.o $(BUILD)/hostemu/obj/pointwise.o $(BUILD)/hostemu/obj/comm.o $(BUILD)/hostemu/obj/psf_prep.o $(BUILD)/hostemu/obj/tiff_io.o $(BUILD)/hostemu/obj/n5_io.o $(BUILD)/hostemu/obj/capi.o

.PHONY: all hostemu clean fft_emu_test
all: $(PKG)/libmvdecon.so

$(PKG)/libmvdecon.so: $(OBJ)
	$(NVCC) -shared -o $@ $(OBJ) -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -lcudart -ldl -lz

$(BUILD)/obj/len_%.o: $(GEN)/len_%.cu $(HDRS)
	@mkdir -p $(BUILD)/obj $(BUILD)/ptxas
	$(NVCC) $(NVFLAGS) -I$(CSRC) -c $< -o $@ 2> $(BUILD)/ptxas/len_$*.log || (cat $(BUILD)/ptxas/len_$*.log; false)
$(BUILD)/obj/registry.o: $(GEN)/registry.cpp $(HDRS)
	@mkdir -p $(BUILD)/obj
	$(NVCC) $(NVFLAGS) -I$(CSRC) -x cu -c $< -o $@ 2> /dev/null
$(BUILD)/obj/engine.o: $(CSRC)/engine.cpp $(HDRS)
	@mkdir -p $(BUILD)/obj
	$(NVCC) $(NVFLAGS) -I$(CSRC) -x cu -c $< -o $@ 2> $(BUILD)/ptxas_engine.log || (cat $(BUILD)/ptxas_engine.log; false)
$(BUILD)/obj/pointwise.o: $(CSRC)/pointwise.cpp $(HDRS)
	@mkdir -p $(BUILD)/obj
	$(NVCC) $(NVFLAGS) -I$(CSRC) -x cu -c $< -o $@ 2> $(BUILD)/ptxas_pointwise.log || (cat $(BUILD)/ptxas_pointwise.log; false)
$(BUILD)/obj/comm.o: $(CSRC)/comm.cpp $(HDRS)
	@mkdir -p $(BUILD)/obj
	$(NVCC) $(NVFLAGS) -I$(CSRC) -x cu -c $< -o $@ 2> $(BUILD)/ptxas_comm.log || (cat $(BUILD)/ptxas_comm.log; false)
$(BUILD)/obj/psf_prep.o: $(CSRC)/psf_prep.cpp $(HDRS)
	@mkdir -p $(BUILD)/obj
	$(NVCC) $(NVFLAGS) -I$(CSRC) -x cu -c $< -o $@ 2> /dev/null
$(BUILD)/obj/tiff_io.o: $(CSRC)/tiff_io.cpp $(HDRS)
	@mkdir -p $(BUILD)/obj
	$(NVCC) $(NVFLAGS) -I$(CSRC) -x cu -c $< -o $@ 2> /dev/null
$(BUILD)/obj/n5_io.o: $(CSRC)/n5_io.cpp $(HDRS)
	@mkdir -p $(BUILD)/obj
	$(NVCC) $(NVFLAGS) -I$(CSRC) -x cu -c $< -o $@ 2> /dev/null
$(BUILD)/obj/capi.o: $(CSRC)/capi.cpp $(HDRS)
	@mkdir -p $(BUILD)/obj
	$(NVCC) $(NVFLAGS) -I$(CSRC) -x cu -c $< -o $@ 2> $(BUILD)/ptxas_capi.log || (cat $(BUILD)/ptxas_capi.log; false)

hostemu: tests/host/libmvdecon_hostemu.so
tests/host/libmvdecon_hostemu.so: $(HOBJ)
	$(CXX) -shared -o $@ $(HOBJ) -lz
$(BUILD)/hostemu/obj/len_%.o: $(GENH)/len_%.cu $(HDRS)
	@mkdir -p $(BUILD)/hostemu/obj
	$(CXX) $(HOSTFLAGS) -x c++ -c $< -o $@
$(BUILD)/hostemu/obj/registry.o: $(GENH)/registry.cpp $(HDRS)
	@mkdir -p $(BUILD)/hostemu/obj
	$(CXX) $(HOSTFLAGS) -x c++ -c $< -o $@
$(BUILD)/hostemu/obj/engine.o: $(CSRC)/engine.cpp $(HDRS)
	@mkdir -p $(BUILD)/hostemu/obj
	$(CXX) $(HOSTFLAGS) -x c++ -c $< -o $@
$(BUILD)/hostemu/obj/pointwise.o: $(CSRC)/pointwise.cpp $(HDRS)
	@mkdir -p $(BUILD)/hostemu/obj
	$(CXX) $(HOSTFLAGS) -x c++ -c $< -o $@
$(BUILD)/hostemu/obj/comm.o: $(CSRC)/comm.cpp $(HDRS)
	@mkdir -p $(BUILD)/hostemu/obj
	$(CXX) $(HOSTFLAGS) -x c++ -c $< -o $@
$(BUILD)/hostemu/obj/psf_prep.o: $(CSRC)/psf_prep.cpp $(HDRS)
	@mkdir -p $(BUILD)/hostemu/obj
	$(CXX) $(HOSTFLAGS) -x c++ -c $< -o $@
$(BUILD)/hostemu/obj/tiff_io.o: $(CSRC)/tiff_io.cpp $(HDRS)
	@mkdir -p $(BUILD)/hostemu/obj
	$(CXX) $(HOSTFLAGS) -x c++ -c $< -o $@
$(BUILD)/hostemu/obj/n5_io.o: $(CSRC)/n5_io.cpp $(HDRS)
	@mkdir -p $(BUILD)/hostemu/obj
	$(CXX) $(HOSTFLAGS) -x c++ -c $< -o $@
$(BUILD)/hostemu/obj/capi.o: $(CSRC)/capi.cpp $(HDRS)
	@mkdir -p $(BUILD)/hostemu/obj
	$(CXX) $(HOSTFLAGS) -x c++ -c $< -o $@

clean:
	rm -rf $(BUILD) $(PKG)/libmvdecon.so tests/host/libmvdecon_hostemu.so

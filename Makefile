# Builds the in-tree shared libraries (sm_100a only; nvcc cross-compiles without a GPU).
NVCC ?= nvcc
ARCH := -gencode arch=compute_100a,code=sm_100a
NVFLAGS := -std=c++17 -O3 -lineinfo $(ARCH) -Xcompiler -fPIC -Xcompiler -Wall -Xcompiler -Wno-unused-function
CSRC := optimet_b200/csrc
OBJS := $(CSRC)/ob_vtac.o $(CSRC)/ob_mie.o $(CSRC)/ob_matvec.o $(CSRC)/ob_pairs.o $(CSRC)/ob_vec.o $(CSRC)/ob_sh.o $(CSRC)/ob_lu.o $(CSRC)/ob_aca.o $(CSRC)/ob_fields.o $(CSRC)/ob_rot.o $(CSRC)/ob_api.o $(CSRC)/ob_multi.o
HDRS := $(wildcard $(CSRC)/*.h $(CSRC)/*.cuh) include/optimet_b200.h
LIB := optimet_b200/liboptimet_b200.so

HOSTLIB := optimet_b200/liboptimet_b200_host.so
HOSTSRC := optimet_b200/host/ob_host.cpp optimet_b200/host/ob_host_capi.cpp
CLI := optimet_b200/optimet3d_b200

all: $(LIB) $(HOSTLIB) $(CLI) oracle

$(CSRC)/%.o: $(CSRC)/%.cu $(HDRS)
	$(NVCC) $(NVFLAGS) -c $< -o $@

$(LIB): $(OBJS)
	$(NVCC) $(ARCH) -shared -o $@ $(OBJS) -ldl -lpthread

$(HOSTLIB): $(HOSTSRC) optimet_b200/host/ob_host.hpp include/optimet_b200.h $(LIB)
	g++ -std=c++11 -O2 -fPIC -shared -Wall $(HOSTSRC) -o $@ -Loptimet_b200 -loptimet_b200 -Wl,-rpath,'$$ORIGIN'

$(CLI): optimet_b200/host/main.cpp $(HOSTLIB)
	g++ -std=c++11 -O2 -Wall optimet_b200/host/main.cpp -o $@ -Loptimet_b200 -loptimet_b200_host -loptimet_b200 -Wl,-rpath,'$$ORIGIN'

oracle:
	$(MAKE) -s -C oracle

clean:
	rm -f $(OBJS) $(LIB) $(HOSTLIB) $(CLI)

.PHONY: all oracle clean

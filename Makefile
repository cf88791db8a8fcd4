# Builds the in-tree shared libraries (sm_100a only; nvcc cross-compiles without a GPU).
NVCC ?= nvcc
ARCH := -gencode arch=compute_100a,code=sm_100a
NVFLAGS := -std=c++17 -O3 -lineinfo $(ARCH) -Xcompiler -fPIC -Xcompiler -Wall -Xcompiler -Wno-unused-function
CSRC := optimet_b200/csrc
OBJS := $(CSRC)/ob_vtac.o $(CSRC)/ob_mie.o $(CSRC)/ob_matvec.o $(CSRC)/ob_vec.o $(CSRC)/ob_sh.o $(CSRC)/ob_api.o
HDRS := $(wildcard $(CSRC)/*.h $(CSRC)/*.cuh) include/optimet_b200.h
LIB := optimet_b200/liboptimet_b200.so

all: $(LIB) oracle

$(CSRC)/%.o: $(CSRC)/%.cu $(HDRS)
	$(NVCC) $(NVFLAGS) -c $< -o $@

$(LIB): $(OBJS)
	$(NVCC) $(ARCH) -shared -o $@ $(OBJS) -ldl

oracle:
	$(MAKE) -s -C oracle

clean:
	rm -f $(OBJS) $(LIB)

.PHONY: all oracle clean

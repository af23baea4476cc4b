# Builds the product library (sm_100a only) and the test-only oracle.  No reference sources are
# compiled here; oracle/Makefile.ref builds the reference's own translation unit into oracle/_ref/.
NVCC      ?= nvcc
CSRC      := gpu_matrix_inversion_b200/csrc
OUT       := gpu_matrix_inversion_b200/libmatinv32.so
NVFLAGS   := -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC,-O3 \
             --expt-relaxed-constexpr -Xptxas -v
CU        := $(wildcard $(CSRC)/*.cu)
OBJ       := $(CU:.cu=.o) $(CSRC)/mat_inv_32.o $(CSRC)/matrix_inversion.o

all: $(OUT) oracle

$(CSRC)/%.o: $(CSRC)/%.cu $(CSRC)/common.cuh $(CSRC)/kernels.h include/matinv_shim.h
	$(NVCC) $(NVFLAGS) -c $< -o $@ 2> $@.ptxas.log || (cat $@.ptxas.log; false)

$(CSRC)/mat_inv_32.o: $(CSRC)/mat_inv_32.cpp include/mat_inv_32.h include/matinv_shim.h
	g++ -O2 -std=c++17 -fPIC -c $< -o $@

$(CSRC)/matrix_inversion.o: $(CSRC)/matrix_inversion.cpp include/matrix_inversion.h include/mat_inv_32.h include/matinv_shim.h
	g++ -O2 -std=c++17 -fPIC -c $< -o $@

$(OUT): $(OBJ)
	$(NVCC) -shared -o $@ $(OBJ) -gencode arch=compute_100a,code=sm_100a -lcudart -ldl

oracle:
	gcc -O3 -mfma -mavx2 -ffp-contract=off -fopenmp -fPIC -shared -o oracle/libgj_oracle.so oracle/gj_oracle.c -lm

# bring-up / regression check of the tcgen05 trailing update without Python (runs in seconds on a fresh GPU box)
tools/tc_check: tools/tc_check.cpp $(OUT) include/matinv_shim.h
	g++ -O2 -std=c++17 $< -Iinclude -I/usr/local/cuda/include -Lgpu_matrix_inversion_b200 -lmatinv32 -L/usr/local/cuda/lib64 -lcudart \
	    -Wl,-rpath,'$$ORIGIN/../gpu_matrix_inversion_b200' -Wl,-rpath,/usr/local/cuda/lib64 -o $@

clean:
	rm -f $(CSRC)/*.o $(CSRC)/*.ptxas.log $(OUT) oracle/libgj_oracle.so

.PHONY: all oracle clean

# oracle/ref.mk -- builds oracle/_ref/libhnh_ref.so from the REFERENCE's own sources where they
# lie (/root/reference; never copied into this repo) against the shims in oracle/shims/.
# TEST INFRASTRUCTURE: the reference's CMake build needs MPI, MKL, Eigen and CombBLAS, none of which
# is in this image (DESIGN.md "oracle"); this recipe compiles the reference's in-tree code
# unmodified and restates only those four third-party interfaces.
REF ?= /root/reference
CXX := /usr/bin/g++
CXXFLAGS := -O3 -march=x86-64-v3 -fopenmp -fPIC -std=c++17 -w -DMKL_ILP64
INC := -Ishims -I$(REF)

all: _ref/libhnh_ref.so

_ref/libhnh_ref.so: ref_driver.cpp shims/hmpi.cpp shims/mpi.h shims/mkl_spblas.h shims/Eigen/Dense shims/CombBLAS/CombBLAS.h
	mkdir -p _ref
	$(CXX) $(CXXFLAGS) $(INC) -shared -o $@ ref_driver.cpp shims/hmpi.cpp \
	    $(REF)/sparse_kernels.cpp $(REF)/common.cpp $(REF)/benchmark_dist.cpp $(REF)/als_conjugate_gradients.cpp

# The same sources with -DHNH_CUDA_PLUGIN: the reference's algorithm classes driving this repo's CUDA kernels through
# the reference's own KernelImplementation interface (needs ../distributed_sddmm_b200/libhnh_b200.so; optional target).
cuda: _ref/libhnh_ref_cuda.so

_ref/libhnh_ref_cuda.so: ref_driver.cpp shims/hmpi.cpp shims/mpi.h shims/mkl_spblas.h shims/Eigen/Dense shims/CombBLAS/CombBLAS.h \
        ../include/hnh/reference_plugin/cuda_kernel.h ../include/hnh_b200.h
	mkdir -p _ref
	$(CXX) $(CXXFLAGS) -DHNH_CUDA_PLUGIN $(INC) -I../include -shared -o $@ ref_driver.cpp shims/hmpi.cpp \
	    $(REF)/sparse_kernels.cpp $(REF)/common.cpp $(REF)/benchmark_dist.cpp $(REF)/als_conjugate_gradients.cpp \
	    -L../distributed_sddmm_b200 -lhnh_b200 -Wl,-rpath,'$$ORIGIN/../../distributed_sddmm_b200'

clean:
	rm -rf _ref

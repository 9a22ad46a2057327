// tests/shim_ref stub: the BLAS TPL singletons (cuBLAS / rocBLAS handles) are not on the sparse path checked here.

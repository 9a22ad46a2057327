// tests/shim_ref stub: KokkosKernels::Impl::Singleton (common/src/KokkosKernels_Singleton.hpp) is only used by the vendor TPL singletons.

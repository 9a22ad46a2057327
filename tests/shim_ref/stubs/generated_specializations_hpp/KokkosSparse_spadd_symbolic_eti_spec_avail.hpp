// tests/shim_ref stub: CMake generates this ETI list in a real build; empty here (no ETI instantiations).

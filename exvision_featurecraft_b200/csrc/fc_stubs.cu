// (no stubs left: every entry point of include/featurecraft_b200.h has a kernel behind it)

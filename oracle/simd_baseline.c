/* placeholder so the Makefile target exists; replaced by the real SIMD baseline */
int azo_simd_placeholder(void) { return 0; }

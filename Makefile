# Top-level build, keeping the reference Makefile's target names and output paths (Makefile:110-111,135-140 there):
#   make shared_library   -> build/bin/stereo_vision_parallel.so   (what stereo_vision/sv.py loads through so_lib_path=)
#   make stereo_vision    -> build/bin/stereo_vision_parallel      (the sequence driver)
# Everything is compiled for sm_100a only by the package's own csrc/Makefile.
PKG := low-cost-hardware-accelerated-vision-based-depth-perception-for-real-time-applications_b200
LIB := $(PKG)/lib/libelas_b200.so

all: shared_library stereo_vision

$(LIB): FORCE
	$(MAKE) -C $(PKG)/csrc

shared_library: $(LIB)
	@mkdir -p build/bin
	cp $(LIB) build/bin/stereo_vision_parallel.so
	ln -sf stereo_vision_parallel.so build/bin/libelas_b200.so

stereo_vision: $(LIB)
	@mkdir -p build/bin
	$(MAKE) -C $(PKG)/csrc driver
	cp $(PKG)/lib/stereo_vision_parallel build/bin/stereo_vision_parallel

clean:
	$(MAKE) -C $(PKG)/csrc clean
	rm -rf build

FORCE:
.PHONY: all shared_library stereo_vision clean FORCE

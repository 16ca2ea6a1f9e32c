# Convenience targets; the native build itself lives in parallel_convolution_b200/build.py.
PY ?= python
.PHONY: build test test-gpu bench reference profiles clean
build:
	$(PY) -m parallel_convolution_b200.build
test: build
	$(PY) -m pytest tests -x -q -m "not gpu"
test-gpu: build
	$(PY) -m pytest tests -x -q -m gpu
bench: build
	$(PY) bench.py
reference:
	$(PY) baseline/build_reference.py
profiles: build
	$(PY) scripts/make_profiles.py
clean:
	rm -rf parallel_convolution_b200/_lib/*.o parallel_convolution_b200/_lib/*.so parallel_convolution_b200/_lib/pconv_native

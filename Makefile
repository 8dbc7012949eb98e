# Convenience targets (the driver uses __graft_entry__.build / pytest / bench.py directly).
PY ?= python

build:
	$(PY) -c "import __graft_entry__ as g; g.build()"

test: build
	$(PY) -m pytest tests -q -m "not gpu"

test-gpu:
	$(PY) -m pytest tests -q -m gpu

bench:
	$(PY) bench.py

reference-arm:
	$(PY) bench.py --impl reference --steps 3 --warmup 1

clean:
	rm -f godot_atmosphere_shader_b200/*.so oracle/liboracle.so tests/cpp/test_node
	rm -rf oracle/_ref tests/hostsim/*.so

.PHONY: build test test-gpu bench reference-arm clean

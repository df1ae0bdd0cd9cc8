# Convenience targets; the driver's entry points are __graft_entry__.build()/smoke() and bench.py.
all: lib oracle
lib:
	$(MAKE) -C rududu_image_codec_b200/csrc
oracle:
	$(MAKE) -C oracle
test:
	python -m pytest tests -q -m "not gpu"
test-gpu:
	python -m pytest tests -q -m gpu
bench:
	python bench.py
clean:
	$(MAKE) -C rududu_image_codec_b200/csrc clean
	$(MAKE) -C oracle clean
.PHONY: all lib oracle test test-gpu bench clean

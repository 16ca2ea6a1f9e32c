"""In-tree native build entry point: ``python setup.py build_ext --inplace`` (or simply
``python -m parallel_convolution_b200.build``) compiles libpconv_host.so with g++ and
libpconv_cuda.so with ``nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo`` into
``parallel_convolution_b200/_lib/``. There is nothing to install: the package is used from the
checkout and loads its shared objects with ctypes."""
import os
import sys

from setuptools import Command, setup

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


class BuildNative(Command):
    description = "build the native host and CUDA (sm_100a) libraries in-tree"
    user_options = [("inplace", "i", "ignored (the build is always in-tree)"), ("force", "f", "rebuild")]

    def initialize_options(self):
        self.inplace = 1
        self.force = 0

    def finalize_options(self):
        pass

    def run(self):
        from parallel_convolution_b200 import build
        build.build_all(force=bool(self.force))


setup(name="parallel_convolution_b200", version="0.1.0", packages=[], cmdclass={"build_ext": BuildNative})

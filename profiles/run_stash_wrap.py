import os, sys, runpy
sys.path.insert(0, os.getcwd())
from vfnerf_b200 import _lib
_lib.LIB_PATH = os.path.abspath("profiles/_build/libvfnerf_prof.so")
sys.argv = ["run_stash_experiment.py"] + sys.argv[1:]
runpy.run_path("profiles/run_stash_experiment.py", run_name="__main__")

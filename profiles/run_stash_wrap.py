"""A/B helper: profiles/run_stash_experiment.py against a differently built library.
Usage: python profiles/run_stash_wrap.py <library.so> [rays] [literal]"""
import os, sys, runpy
sys.path.insert(0, os.getcwd())
from vfnerf_b200 import _lib
_lib.LIB_PATH = os.path.abspath(sys.argv[1])
sys.argv = ["run_stash_experiment.py"] + sys.argv[2:]
runpy.run_path("profiles/run_stash_experiment.py", run_name="__main__")

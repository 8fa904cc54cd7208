import sys
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
from _scenes import *
from sosba_loader import load_package
pkg=load_package(); lib=pkg.load()
sc=scene(**CONFIG_B)
h=open_handle(lib,sc); P,k=upload(h,sc)
h.reset_oob(); h.linearize_all(False); h.apply_res()
for i in range(4): h.solve_system()

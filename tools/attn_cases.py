import os, sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tools")
import torch
import dev_attn as da
for args in [(1,128,256,1,40),(1,256,128,1,40),(1,128,192,1,40),(1,128,192,1,24),(1,128,192,1,16),(1,128,192,1,32),(1,64,192,1,40),(1,128,192,2,40),(1,128,192,8,40)]:
    da.check(*args)

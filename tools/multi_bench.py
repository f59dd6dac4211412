import sys, os, time
sys.path.insert(0,'.'); sys.path.insert(0,'tests'); sys.path.insert(0,'tools')
import numpy as np, torch
import blingfire_b200 as bf
import corpus
from _common import model_path
torch.zeros(1, device='cuda')
for name, pool in (("bert_base_tok.bin","MULTI"),("bert_multi_cased.bin","MULTI"),("bert_base_tok.bin","EN")):
    h=bf.load_model(model_path(name))
    text, offs = corpus.gen_docs(pool, 400000, seed=4, fixed_len=512, emoji_every=16 if pool=="MULTI" else 0)
    n=len(offs)-1; nbytes=int(offs[-1])
    d_text=torch.empty(nbytes+64,dtype=torch.uint8,device='cuda'); d_text[:nbytes].copy_(torch.from_numpy(text))
    d_offs=torch.from_numpy(offs).cuda(); d_ids=torch.empty((n,512),dtype=torch.int32,device='cuda'); d_counts=torch.zeros(n,dtype=torch.int32,device='cuda')
    st=torch.cuda.current_stream()
    def step(): bf.text_to_ids_batch_device(h, d_text.data_ptr(), d_offs.data_ptr(), n, nbytes, d_ids.data_ptr(), d_counts.data_ptr(), 512, 100, st.cuda_stream, max_doc_bytes=int(np.diff(offs).max()))
    for _ in range(3): step()
    torch.cuda.synchronize(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record(); 
    for _ in range(10): step()
    e1.record(); torch.cuda.synchronize()
    ms=e0.elapsed_time(e1)/10
    print(name, pool, '%.1f GB/s (%.2f ms, %d tokens)'%(nbytes/ms/1e6, ms, int(d_counts.sum())))
    bf.free_model(h)

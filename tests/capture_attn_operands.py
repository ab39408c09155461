import sys, math, torch, numpy as np
sys.path.insert(0, '/root/repo')
from oracle import rmem_oracle as O
H, W, N_OBJ = 481, 849, 10
torch.set_num_threads(8)
sd = O.make_state_dict("r50_deaotl", seed=0, sharpen=4.0)
frames = O.synthetic_frames(9, H, W, seed=1000)
label0 = O.synthetic_label(H, W, N_OBJ)
eng = O.OracleEngine(sd, O.OracleConfig(former_mem_len=1, latter_mem_len=7), long_term_mem_gap=1)
caps = []
orig = O.long_term_attention
def hook(q, kb, vb, scale_dim, n_head=1):
    if hook.on and kb.shape[0] > 1:
        caps.append((q.clone(), kb.clone()))
    return orig(q, kb, vb, scale_dim, n_head)
hook.on = False
O.long_term_attention = hook
with torch.no_grad():
    eng.add_reference_frame(frames[0:1], label0, obj_nums=[N_OBJ], frame_step=0)
    for i in range(12):
        if i == 11: hook.on = True
        lg = eng.match_propogate_one_frame(frames[1 + i % 8: 2 + i % 8], output_size=(H, W))
        eng.update_memory(O.logits_to_label(lg))
        print(i, len(eng.aot_engines[0].long_memories_indexes), flush=True)
for l, (q, kb) in enumerate(caps):
    np.savez(f'scratch/attn_cap_layer{l}.npz', q=q.numpy(), k=kb.numpy())
    s = (q @ kb.flatten(0,1).t()) / math.sqrt(128) * 1.4426950408889634
    print('layer', l, 'T', kb.shape[0], 'score log2 units: std %.2f rowmax mean %.2f, max %.2f min %.2f' % (s.std(), s.max(1).values.mean(), s.max(), s.min()))

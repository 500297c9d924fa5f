"""which python lines still force a device read (Tensor.__bool__ / item) in a plugin-elided forward of the reference's modules"""
import sys, traceback, collections, os
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(R, "oracle", "refshim")); sys.path.insert(0, R)
import torch
import load_reference
ref = load_reference.load_full()
from dmx_compressor_b200 import opt, plugin, elide
cfg = dict(vocab_size=512, max_position_embeddings=128, hidden_size=128, num_hidden_layers=1, ffn_dim=256, num_attention_heads=2, dropout=0.0)
torch.manual_seed(0)
net = opt.OPTStack(cfg, mods=ref.nn).cuda().to(torch.bfloat16).eval()
for m in net.modules():
    if isinstance(m, ref.nn.DmxModule):
        for rule in ref.config_rules.BASIC:
            if isinstance(m, rule.module_types):
                m.configure(rule.module_config); break
ids = torch.randint(0, 512, (2, 64)).cuda()
sites = collections.Counter()
orig_bool, orig_item = torch.Tensor.__bool__, torch.Tensor.item
def site():
    st = traceback.extract_stack(limit=4)
    return " <- ".join(f"{f.filename.split('/')[-1]}:{f.lineno}" for f in reversed(st[:-1]))
def hb(self):
    sites["bool " + site()] += 1
    return orig_bool(self)
def hi(self):
    sites["item " + site()] += 1
    return orig_item(self)
plugin.install(elide=True)
with torch.no_grad(), elide.enabled():
    net(ids); net(ids)
    torch.Tensor.__bool__, torch.Tensor.item = hb, hi
    net(ids)
    torch.Tensor.__bool__, torch.Tensor.item = orig_bool, orig_item
plugin.uninstall()
for k, v in sites.most_common():
    print(v, k)

import sys, json, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/oracle")
import bcos_b200, bcos_oracle as OR
from bcos_b200.engine import CLIPViTPlan
from bcos_b200.utils import synth
sd = synth.synth_state_dict(OR.clip_vit_state_shapes(224, 32, 768, 12, 512), 0)
g = torch.Generator().manual_seed(0)
t = torch.nn.functional.normalize(torch.randn(512, generator=g), dim=0)
for mode in ("parity", "throughput"):
    B = 512
    plan = CLIPViTPlan(sd, B, mode=mode, device="cuda", input_u8=True)
    x = torch.from_numpy(synth.synth_images_u8(32, 224, 3)).repeat(B // 32, 1, 1, 1).cuda()
    plan.load_input(x); plan.capture()
    for _ in range(2): plan.explain_direction(None, t)
    torch.cuda.synchronize()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e[0].record()
    for _ in range(5): plan.embed(None)
    e[1].record()
    for _ in range(5): plan.explain_direction(None, t)
    e[2].record(); torch.cuda.synchronize()
    print(json.dumps({"net": "CLIP ViT-B/32", "mode": mode, "batch": B, "embed_ms": round(e[0].elapsed_time(e[1]) / 5, 3), "embed_explain_ms": round(e[1].elapsed_time(e[2]) / 5, 3),
                      "embed_explain_img_s": round(B / (e[1].elapsed_time(e[2]) / 5) * 1e3, 1), "launches": plan.num_launches()}), flush=True)
    del plan; torch.cuda.empty_cache()

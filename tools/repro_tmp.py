import os, sys, subprocess
sys.path.insert(0, "/root/repo/r-super_b200")
if len(sys.argv) > 1:
    import torch
    from rsuper_b200 import ops
    N,D,H,W,Cin,Cout,pz = [int(v) for v in sys.argv[1:8]]
    dev="cuda"
    g=torch.Generator().manual_seed(0)
    x=torch.randn(N,D,H,W,Cin,generator=g).to(dev).to(torch.bfloat16)
    w=(torch.randn(Cout,Cin,3,3,3,generator=g)/70).to(dev)
    y=torch.zeros(N,D,H,W,Cout,dtype=torch.bfloat16,device=dev)
    ost=torch.zeros(N,Cout,2,device=dev)
    wp=ops.conv3_pack_weights(w)
    for i in range(6):
        ops.conv3_forward(x,wp,y,out_stats=ost,planes_per_item=pz)
        torch.cuda.synchronize()
    print("ok", y.float().abs().mean().item())
else:
    for cfg in ["2 64 64 64 128 192 2", "2 64 64 64 192 128 1", "2 64 64 64 192 128 2", "2 32 32 32 128 128 2", "2 64 64 64 64 128 2", "2 64 64 64 32 96 2", "2 64 64 64 32 80 2", "2 64 64 64 32 112 2"]:
        r = subprocess.run([sys.executable, __file__] + cfg.split(), capture_output=True, text=True)
        print(cfg, "->", (r.stdout.strip().splitlines() or ["FAIL"])[-1][:60], "| rc", r.returncode, flush=True)

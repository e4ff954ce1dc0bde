"""Recipe of the NaViT config-5-geometry parity case (shared by make_golden.py, which runs the UNMODIFIED reference on
it in the build container, and by the tests, which rebuild the same weights and images from the seeds)."""
import torch

# BASELINE.json configs[4] GEOMETRY (dim 1024, depth 6, heads 16, mlp 4096).  The weights (76 M parameters) are not
# stored: the drop-in's constructor consumes the RNG exactly like the reference's (tests/test_navit.py), so the test
# rebuilds them from `seed` with navit_config5_model() below; stored are the image sizes, the reference's fp32
# logits and the reference's own bf16 logits (its noise floor) on the same bf16-representable weights and inputs.
NAVIT_C5 = dict(seed=7, input_seed=107, gamma_seed=1007,
                kwargs=dict(image_size=512, patch_size=16, num_classes=1000, dim=1024, depth=6, heads=16, mlp_dim=4096),
                # 32x32 patches = 1024 tokens (8 key blocks of 128), a 1-token image, ragged sizes around the 64 / 128
                # key-block edges, one wide and one tall strip
                sizes=[(512, 512), (16, 16), (128, 144), (272, 240), (64, 512), (512, 32), (208, 160), (336, 496),
                       (16, 32), (144, 224)])


def navit_config5_model(cls, spec=NAVIT_C5):
    """`cls` = the reference's NaViT (here) or the drop-in's (tests): identical parameters from the same seeds."""
    torch.manual_seed(spec["seed"])
    model = cls(**spec["kwargs"]).eval()
    g = torch.Generator().manual_seed(spec["gamma_seed"])
    with torch.no_grad():
        for n, p in model.named_parameters():
            if n.endswith("gamma"):
                p.add_(0.1 * torch.randn(p.shape, generator=g))
            p.copy_(p.bfloat16().float())
    return model


def navit_config5_images(spec=NAVIT_C5):
    torch.manual_seed(spec["input_seed"])
    return [torch.randn(3, h, w).bfloat16() for h, w in spec["sizes"]]

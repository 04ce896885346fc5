"""The drop-in boundary, proven with the reference's own call sequence.

CPU part (this file, ``-m "not gpu"``): every test runs in a fresh interpreter, because ``install()``
edits ``sys.modules``.
  * stand-alone install (no reference on sys.path): the two submodule names resolve to the mirror;
  * with the reference tree importable (build container only; skipped elsewhere): after ``install()``
    the reference's ``curriculums`` imports (curriculums.py:1 needs the reference's own
    ``generators.neural_rendering``), ``extract_metadata`` works, the train script's class lookups
    (train_double_latent_semantic.py:20-22, 116, 142) find the mirror classes;
  * a generator built from the REFERENCE classes and saved with ``torch.save(generator)``
    (train_double_latent_semantic.py:128-150 / render_multiview_images_double_semantic.py:58) loads
    under the mirror with an identical state_dict and the same parameter order, so torch_ema's
    positional ``copy_to`` / ``restore`` (``param.data.copy_``) lands on the right tensors.
GPU part: tests/test_gpu_parity.py::test_render_script_call_sequence replays
render_multiview_images_double_semantic.py:43-65 against a golden.
"""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("FENERF_REFERENCE_ROOT", "/root/reference")
HAVE_REF = os.path.isdir(os.path.join(REF, "generators"))
needs_ref = pytest.mark.skipif(not HAVE_REF, reason="reference tree not present (build container only)")


def _run(code, *args, timeout=600):
    env = dict(os.environ, PYTHONDONTWRITEBYTECODE="1")
    r = subprocess.run([sys.executable, "-c", code, ROOT, REF] + list(args), capture_output=True, text=True,
                       timeout=timeout, env=env)
    assert r.returncode == 0, "child failed:\n%s\n%s" % (r.stdout[-3000:], r.stderr[-3000:])
    return r.stdout


# the four dead imports of the reference that this image cannot satisfy; a user's
# environment has the real packages
_STUBS = r'''
import sys, types
def _stub(name, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items(): setattr(m, k, v)
    sys.modules[name] = m
    return m
mpl = _stub("matplotlib"); mpl.pyplot = _stub("matplotlib.pyplot")
import numpy.lib
sys.modules["numpy.lib"].type_check = _stub("numpy.lib.type_check", imag=None)
_stub("fid_evaluation", output_images=None)
k = _stub("kornia"); k.filters = _stub("kornia.filters", filter2D=None)
'''


def test_standalone_install_resolves_the_two_submodules():
    out = _run(r'''
import sys
sys.path.insert(0, sys.argv[1])
import fenerf_b200
g, s = fenerf_b200.install()
from generators import generators
from siren import siren
assert generators is g and siren is s
import generators.generators as gg, siren.siren as ss
assert gg is g and ss is s
cls = getattr(generators, "DoubleImplicitGenerator3d")
assert cls.__module__ == "generators.generators", cls.__module__
assert hasattr(siren, "TextureEmbeddingPiGAN256SEMANTICDISENTANGLE_DIM_96") and hasattr(siren, "TALLSIREN")
fenerf_b200.install()      # idempotent
print("ok")
''')
    assert "ok" in out


@needs_ref
def test_install_keeps_the_reference_packages_importable():
    out = _run(_STUBS + r'''
sys.path.insert(0, sys.argv[2]); sys.path.insert(0, sys.argv[1])
import fenerf_b200
fenerf_b200.install()
import curriculums                                    # curriculums.py:1 -> generators.neural_rendering (reference's)
import generators.neural_rendering as nr
assert nr.__file__.startswith(sys.argv[2]), nr.__file__
from generators import generators                     # train_double_latent_semantic.py:20
from siren import siren                               # :22
assert generators.__name__ == "fenerf_b200.generators.generators", generators.__name__
assert siren.__name__ == "fenerf_b200.siren.siren"
import importlib.util                                 # reference-only subpackages still reachable (networks.py:18)
assert importlib.util.find_spec("siren.op") is not None and importlib.util.find_spec("generators.BiSeNet") is not None
md = curriculums.extract_metadata(curriculums.CelebA_double_semantic_texture_embedding_256_dim_96, 0)
SIREN = getattr(siren, md['model'])                   # :116
gen_cls = getattr(generators, md['generator'])        # :142
assert gen_cls.__module__ == "generators.generators" and SIREN.__module__ == "siren.siren"
md2 = curriculums.extract_metadata(curriculums.CelebA, 0)
assert hasattr(siren, md2['model']) and hasattr(generators, md2['generator'])
md3 = curriculums.extract_metadata(curriculums.CelebA_double_semantic, 0)
assert hasattr(siren, md3['model']) and hasattr(generators, md3['generator'])
print("ok", md['model'], md['generator'])
''')
    assert "ok TextureEmbeddingPiGAN256SEMANTICDISENTANGLE_DIM_96 DoubleImplicitGenerator3d" in out


_SAVE_WITH_REFERENCE = _STUBS + r'''
sys.path.insert(0, sys.argv[2])
import warnings; warnings.simplefilter("ignore")
import torch
from generators import generators
from siren import siren
assert generators.__file__.startswith(sys.argv[2])
torch.manual_seed(0)
model = sys.argv[4]
if model == "A":
    gen = generators.ImplicitGenerator3d(siren.TALLSIREN, 256, 4)
else:
    gen = generators.DoubleImplicitGenerator3d(siren.SIRENBASELINESEMANTICDISENTANGLE, 256, 256, 22)
gen.set_device("cpu")
gen.step, gen.epoch = 1234, 7
torch.save(gen, sys.argv[3])                                        # train_double_latent_semantic.py:523 style
names = [n for n, _ in gen.named_parameters()]
torch.save({"names": names, "state": gen.state_dict()}, sys.argv[3] + ".meta")
'''

_LOAD_WITH_MIRROR = r'''
import sys
sys.path.insert(0, sys.argv[1])
import torch, fenerf_b200
fenerf_b200.install()
gen = torch.load(sys.argv[3], map_location="cpu", weights_only=False)   # render_multiview_images_double_semantic.py:58
meta = torch.load(sys.argv[3] + ".meta", map_location="cpu", weights_only=False)
assert type(gen).__module__ == "generators.generators", type(gen).__module__
assert type(gen).__mro__[1].__name__ == "_RenderSkeleton", "not the mirror class"
assert gen.step == 1234 and gen.epoch == 7 and gen.device == "cpu"
# parameter registration order (torch_ema is positional) and state_dict equality
assert [n for n, _ in gen.named_parameters()] == meta["names"]
sd = gen.state_dict()
assert list(sd.keys()) == list(meta["state"].keys())
for k, v in meta["state"].items():
    assert torch.equal(sd[k], v), k
# what ExponentialMovingAverage.copy_to does: param.data.copy_(shadow) in parameters() order
shadow = [p.detach().clone() + 0.125 for p in gen.parameters()]
for s_param, param in zip(shadow, gen.parameters()):
    param.data.copy_(s_param.data)
for (n, p), s in zip(gen.named_parameters(), shadow):
    assert torch.equal(p, s), n
# the sequence of render_multiview_images_double_semantic.py:59-65 up to the render call
gen.softmax_label = False
gen.neural_renderer_img = None
gen.neural_renderer_seg = None
gen.set_device("cpu")            # runs generate_avg_frequencies with the mapping network on the CPU
gen.eval()
assert hasattr(gen, "avg_frequencies") or hasattr(gen, "avg_frequencies_geo")
# and a checkpoint saved under the mirror carries the reference's module paths
import io, pickletools
buf = io.BytesIO(); torch.save(gen, buf)
assert b"generators.generators" in buf.getvalue() and b"fenerf_b200" not in buf.getvalue()
print("ok")
'''


@needs_ref
@pytest.mark.parametrize("model", ["A", "D"])
def test_reference_pickle_loads_under_the_mirror(tmp_path, model):
    path = str(tmp_path / "generator.pth")
    _run(_SAVE_WITH_REFERENCE, path, model)
    out = _run(_LOAD_WITH_MIRROR, path)
    assert "ok" in out


@needs_ref
def test_mirror_pickle_loads_under_the_reference(tmp_path):
    """The other direction: a whole-module checkpoint written under this library is a valid
    reference checkpoint (same module paths, attribute names and state_dict)."""
    path = str(tmp_path / "generator.pth")
    _run(r'''
import sys
sys.path.insert(0, sys.argv[1])
import torch, fenerf_b200
g, s = fenerf_b200.install()
torch.manual_seed(0)
gen = g.ImplicitGenerator3d(s.TALLSIREN, 256, 4)
gen.set_device("cpu")
torch.save(gen, sys.argv[3])
torch.save(gen.state_dict(), sys.argv[3] + ".sd")
''', path)
    out = _run(_STUBS + r'''
sys.path.insert(0, sys.argv[2])
import warnings; warnings.simplefilter("ignore")
import torch
gen = torch.load(sys.argv[3], map_location="cpu", weights_only=False)
import generators.generators as gg
assert gg.__file__.startswith(sys.argv[2]) and type(gen) is gg.ImplicitGenerator3d
sd = torch.load(sys.argv[3] + ".sd", map_location="cpu")
for k, v in gen.state_dict().items():
    assert torch.equal(v, sd[k]), k
torch.manual_seed(3)
with torch.no_grad():
    px, poses = gen(torch.randn(1, 256), img_size=8, fov=12, ray_start=0.88, ray_end=1.12, num_steps=4, h_stddev=0.,
                    v_stddev=0., h_mean=1.57, v_mean=1.57, hierarchical_sample=True, clamp_mode='relu', nerf_noise=0.)
assert px.shape == (1, 3, 8, 8)
print("ok")
''', path)
    assert "ok" in out

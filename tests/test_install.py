"""semiuhpe_b200.install (INTEGRATION.md option B): module injection in the documented order -- patch first,
import the reference's agent afterwards -- against a minimal fake ``src`` tree, and against the real reference
tree when it is present (this container; the GPU box has no /root/reference).  No kernel runs: the checks
are about which function objects the reference ends up bound to."""
import os
import subprocess
import sys
import textwrap

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFERENCE = "/root/reference"

FAKE_AGENT = '''
    from src import utils
    from src.fisher.fisher_utils import vmf_loss as fisher_NLL
    from src.fisher.fisher_utils import fisher_CE, batch_torch_A_to_R, fisher_entropy
    from src.laplace.rotation_laplace import NLL_loss as laplace_NLL
    from src.utils import compute_euler_angles_from_rotation_matrices

    class SSLAgent:
        def __init__(self, config):
            self.config = config
        def compute_dynamic_entropy_threshold(self, ulb_train_bar):
            return "reference method"
        @staticmethod
        def compute_err_deg_from_quats(pred, gt):
            return "reference quats"
        @staticmethod
        def compute_err_deg_from_matrices(pred, gt, gt_euler=None):
            return "reference matrices"
'''

CHECK = '''
    import sys, types
    sys.path.insert(0, {root!r})
    sys.path.insert(0, {tree!r})
    {stubs}
    from semiuhpe_b200 import install
    done = install.patch_reference()                 # BEFORE the agent is imported (the documented order)
    assert "src.agent" not in sys.modules
    assert any(d.startswith("src.agent.SSLAgent") and d.endswith("(on import)") for d in done), done
    import src.agent as ref_agent                    # what `import train` does
    import semiuhpe_b200.fisher.fisher_utils as ours_f
    import semiuhpe_b200.laplace.rotation_laplace as ours_l
    import semiuhpe_b200.utils as ours_u
    import semiuhpe_b200.agent as ours_a
    assert ref_agent.fisher_NLL is ours_f.vmf_loss and ref_agent.fisher_CE is ours_f.fisher_CE
    assert ref_agent.fisher_entropy is ours_f.fisher_entropy and ref_agent.batch_torch_A_to_R is ours_f.batch_torch_A_to_R
    assert ref_agent.laplace_NLL is ours_l.NLL_loss
    assert ref_agent.compute_euler_angles_from_rotation_matrices is ours_u.compute_euler_angles_from_rotation_matrices
    assert sys.modules["src.utils"].compute_euler_angles_from_rotation_matrices is ours_u.compute_euler_angles_from_rotation_matrices
    cls = ref_agent.SSLAgent
    assert cls.compute_dynamic_entropy_threshold is ours_a.compute_dynamic_entropy_threshold
    assert cls.compute_err_deg_from_matrices is ours_a.compute_err_deg_from_matrices
    assert cls.compute_err_deg_from_quats is ours_a.compute_err_deg_from_quats
    assert cls._reference_compute_dynamic_entropy_threshold.__module__ == "src.agent"
    # config.save_feat: handed back to the reference's own method instead of silently skipped
    class Cfg: save_feat = True; left_ratio = 0.95
    agent = object.__new__(cls); agent.config = Cfg()
    called = []
    cls._reference_compute_dynamic_entropy_threshold = lambda self, bar: called.append(bar) or "delegated"
    assert agent.compute_dynamic_entropy_threshold("loader") == "delegated" and called == ["loader"]
    # a second patch_reference() is idempotent and now patches in place
    done2 = install.patch_reference()
    assert "src.agent.SSLAgent" in done2 and cls.compute_dynamic_entropy_threshold is ours_a.compute_dynamic_entropy_threshold
    print("INSTALL-OK")
'''


def _run(tree, stubs=""):
    code = textwrap.dedent(CHECK).format(root=ROOT, tree=tree, stubs="{stubs}").replace("{stubs}", stubs)
    proc = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert proc.returncode == 0 and "INSTALL-OK" in proc.stdout, proc.stdout + proc.stderr


def test_patch_before_import_on_a_fake_tree(built, tmp_path):
    src = tmp_path / "src"
    (src / "fisher").mkdir(parents=True)
    (src / "laplace").mkdir()
    for d in (src, src / "fisher", src / "laplace"):
        (d / "__init__.py").write_text("")
    (src / "fisher" / "fisher_utils.py").write_text("def vmf_loss(*a): return 'ref'\nfisher_CE = batch_torch_A_to_R = fisher_entropy = vmf_loss\n")
    (src / "laplace" / "rotation_laplace.py").write_text("def NLL_loss(*a): return 'ref'\n")
    (src / "utils.py").write_text("def compute_euler_angles_from_rotation_matrices(*a, **k): return 'ref'\nclass TrainClock: pass\n")
    (src / "agent.py").write_text(textwrap.dedent(FAKE_AGENT))
    _run(str(tmp_path))


@pytest.mark.skipif(not os.path.isdir(os.path.join(REFERENCE, "src")), reason="reference tree not present on this host")
def test_patch_before_import_on_the_real_reference(built):
    """The unmodified src/agent.py of the reference; its heavy dependencies that this image lacks (pytorch3d, the
    backbone zoo behind src.networks) are stubbed -- they are not on the path."""
    stubs = "; ".join([
        'p3d = types.ModuleType("pytorch3d")', 'p3d.transforms = types.ModuleType("pytorch3d.transforms")',
        'sys.modules["pytorch3d"] = p3d', 'sys.modules["pytorch3d.transforms"] = p3d.transforms',
        'nets = types.ModuleType("src.networks")', 'nets.get_network = lambda config: None',
        'sys.modules["src.networks"] = nets'])
    _run(REFERENCE, stubs)


def test_save_feat_without_a_reference_method_is_an_error(built):
    from semiuhpe_b200.agent import compute_dynamic_entropy_threshold

    class Cfg:
        save_feat = True
        left_ratio = 0.95

    class Bare:
        config = Cfg()

    with pytest.raises(NotImplementedError, match="save_feat"):
        compute_dynamic_entropy_threshold(Bare(), [])

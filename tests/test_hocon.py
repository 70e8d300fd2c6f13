"""The flat-HOCON reader against an inline sample and, where the reference tree is mounted, its real configs."""
import glob
import os

import pytest

from ood_dfq_b200 import hocon

SAMPLE = '''
#  ------------ General options ----
model_name = resnet18
generateDataPath = "./data/x_beta0.1_gamma0.5_group"   # trailing comment
dataset = "imagenet" # options: imagenet | cifar100
nThreads = 8  // c-style comment
batchSize = 64 # 4  # batchsize
weightDecay = 1e-4
lr_S = 0.000001
step_S = [100,200,350]
empty = []
flag = true
url = "http://host/path#frag"
qw = 3
qa = 3
'''


def test_sample():
    c = hocon.loads(SAMPLE)
    assert c["model_name"] == "resnet18" and c["dataset"] == "imagenet" and c["nThreads"] == 8
    assert c["batchSize"] == 64 and c["weightDecay"] == 1e-4 and c["lr_S"] == 1e-6
    assert c["step_S"] == [100, 200, 350] and c["empty"] == [] and c["flag"] is True
    assert c["url"] == "http://host/path#frag"
    assert c["generateDataPath"] == "./data/x_beta0.1_gamma0.5_group"


@pytest.mark.parametrize("bad", ["a = [1,\n2]", "a = { b = 1 }", "just words"])
def test_rejects_what_it_does_not_understand(bad):
    with pytest.raises(ValueError):
        hocon.loads(bad)


REF = "/root/reference/config"


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not mounted (GPU box)")
def test_reads_every_reference_config():
    files = sorted(glob.glob(os.path.join(REF, "*.hocon")))
    assert len(files) >= 10
    for f in files:
        s = hocon.QuantSettings.from_file(f)
        assert s.qw in (2, 3, 4, 5, 6, 8) and s.qa in (2, 3, 4, 5, 6, 8) and s.img_size in (28, 32, 224)
        assert s.lam == 1000.0 and s.eps == 0.01                  # options.py:64-65 overrides
    s = hocon.QuantSettings.from_file(os.path.join(REF, "pathmnist_resnet18_w2a2.hocon"))
    assert (s.qw, s.qa, s.batchSize, s.img_size, s.nClasses) == (2, 2, 64, 28, 9)
    s = hocon.QuantSettings.from_file(os.path.join(REF, "imagenet.hocon"))
    assert (s.model_name, s.batchSize, s.qw, s.temperature, s.alpha) == ("resnet18", 4, 3, 20.0, 20.0)

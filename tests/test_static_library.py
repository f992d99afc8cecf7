"""What can be said about liborb_b200.so without a GPU: it holds device code for sm_100a only, the kernels DESIGN.md
names are in it, and the instruction mix matches the design (vectorised 128-bit loads and cp.async in the streaming
kernels, bulk copies + mbarrier only in the cooperative partition, ballots in the count kernels, no spilled partition)."""
import re
import shutil
import subprocess

import pytest

pytestmark = pytest.mark.skipif(shutil.which("cuobjdump") is None, reason="cuobjdump not on PATH")


@pytest.fixture(scope="module")
def sass(orb):
    out = subprocess.run(["cuobjdump", "-sass", str(orb.LIB_PATH)], capture_output=True, text=True, check=True).stdout
    funcs, cur = {}, None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            funcs[cur] = []
        elif cur is not None:
            funcs[cur].append(line)
    return {k: "\n".join(v) for k, v in funcs.items()}


def kernels(sass, needle):
    return {k: v for k, v in sass.items() if needle in k}


def test_device_code_is_sm_100a_only(orb):
    out = subprocess.run(["cuobjdump", "--list-elf", str(orb.LIB_PATH)], capture_output=True, text=True, check=True).stdout
    archs = set(re.findall(r"\.(sm_\w+)\.cubin", out))
    assert archs == {"sm_100a"}, archs
    ptx = subprocess.run(["cuobjdump", "--list-ptx", str(orb.LIB_PATH)], capture_output=True, text=True).stdout
    assert "sm_9" not in ptx and "sm_8" not in ptx          # no multi-arch fallbacks


def test_hot_path_kernels_are_present(sass):
    for name in ("k_count_stream", "k_count_cells", "k_update", "k_level_persistent", "k_split", "k_partition_coop", "k_partition_cells",
                 "k_partition_warp", "k_hoare_scan", "k_hoare_swap", "k_bbox", "k_sel_stream", "k_sel_resolve", "k_sel_percell",
                 "k_sel_finish", "k_sel_fine", "k_sel_fin_a", "k_sel_gather", "k_sel_fin_b", "k_selx_resolve", "k_xd_compact", "k_xf_finish_block"):
        assert kernels(sass, name), f"{name} is not in the library"


def test_instruction_mix_follows_the_design(sass):
    for name, body in kernels(sass, "k_count_stream").items():
        assert "LDGSTS" in body or "LDG.E.128" in body, name            # cp.async ring / 128-bit loads of the column
        assert "VOTE" in body and "ATOMS" in body, name                  # ballots, shared-memory accumulation per cell
    coop = kernels(sass, "k_partition_coop")
    assert len(coop) == 1
    body = next(iter(coop.values()))
    assert "UBLKCP" in body and "SYNCS" in body                          # bulk-copy tile loads + mbarrier (ORB_PART_BULK=1 path)
    assert "LDGSTS" in body                                              # the default cp.async ring
    others = {k: v for k, v in sass.items() if "k_partition_coop" not in k}
    assert not any("UBLKCP" in v for v in others.values())
    for name, body in kernels(sass, "k_bbox").items():
        if "k_bbox_" not in name and "k_apply" not in name:
            assert "LDG.E.128" in body, name                             # float4 loads of the three columns
    for name, body in sass.items():
        assert "HMMA" not in body and "UTCHMMA" not in body and "IMMA" not in body, name    # no tensor-core work on this path

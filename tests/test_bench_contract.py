"""bench.py's driver contract on the CPU-runnable arm (`--impl reference` times the unmodified
reference package from baseline/_ref -- the oracle port only where that is absent -- on the host
cores): exactly ONE line on stdout, valid JSON, the keys the driver reads."""
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]


def test_reference_arm_prints_one_json_line():
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600, cwd=str(ROOT))
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout[:500]
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "Gpixel/s" and d["higher_is_better"] is True
    assert d["metric"].startswith("waterfall Gpixel/s") and d["value"] > 0 and d["steps"] == 1
    have_ref = (ROOT / "baseline" / "_ref" / "rfi_toolbox" / "__init__.py").exists()
    assert d["cpu_baseline"]["kind"] == ("reference" if have_ref else "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    assert "configs[1]" in d["config"]["workload"] and d["vs_baseline"] is None

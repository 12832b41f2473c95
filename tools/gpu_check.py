"""Run the GPU test files one pytest node group at a time, each in its own process with a timeout, so one
faulting or hanging kernel cannot take the rest of a gpurun call with it.  Writes gpurun_out/check_*.log and
a one-line-per-group summary gpurun_out/check_summary.txt.

    python tools/gpu_check.py [group ...]
"""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
GROUPS = {
    "gemm_exact": "tests/test_gpu_kernels.py::test_gemm_fp32_out_exact_small_integers",
    "gemm_pw": "tests/test_gpu_kernels.py::test_gemm_pointwise",
    "gemm_batched": "tests/test_gpu_kernels.py::test_gemm_batched_weights",
    "conv3x3": "tests/test_gpu_kernels.py::test_gemm_conv3x3",
    "dw_plain": "tests/test_gpu_kernels.py::test_dwconv_plain",
    "dw_gate": "tests/test_gpu_kernels.py::test_dwconv_gate",
    "pwdw": "tests/test_gpu_kernels.py::test_pwdw",
    "mdta": "tests/test_gpu_kernels.py::test_mdta",
    "prompt": "tests/test_gpu_kernels.py::test_prompt_gen",
    "patch": "tests/test_gpu_kernels.py::test_patch_embed",
    "blend": "tests/test_gpu_kernels.py::test_tile_blend_matches_reference_loop",
    "ops": "tests/test_gpu_model.py::test_engine_op_by_op",
    "forward": "tests/test_gpu_model.py::test_forward_matches_reference_golden",
    "model_misc": "tests/test_gpu_model.py -k 'deterministic or biasfree or loud'",
}


def main():
    os.makedirs(OUT, exist_ok=True)
    names = sys.argv[1:] or list(GROUPS)
    summary = []
    for n in names:
        t = time.time()
        cmd = f"{sys.executable} -m pytest {GROUPS[n]} -m gpu -q -x --no-header -p no:cacheprovider -s"
        log = os.path.join(OUT, f"check_{n}.log")
        try:
            with open(log, "w") as f:
                rc = subprocess.run(cmd, shell=True, cwd=ROOT, stdout=f, stderr=subprocess.STDOUT, timeout=420).returncode
        except subprocess.TimeoutExpired:
            rc = "TIMEOUT"
        tail = ""
        try:
            lines = open(log).read().strip().splitlines()
            tail = lines[-1] if lines else ""
        except OSError:
            pass
        line = f"{n:14s} rc={rc} {time.time() - t:6.1f}s  {tail}"
        print(line, flush=True)
        summary.append(line)
        with open(os.path.join(OUT, "check_summary.txt"), "w") as f:
            f.write("\n".join(summary) + "\n")
        if rc == "TIMEOUT":
            # a hung kernel may have wedged the device: stop here rather than pile more work on it
            print("stopping after timeout", flush=True)
            break


if __name__ == "__main__":
    main()

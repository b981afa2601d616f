import sys, os
sys.path.insert(0, os.path.join(os.getcwd(), "scripts"))
import scale_bench as sb
sb.CASES["a_260x352_b64"] = (260, 352, 64, 50000)
sb.CASES["a_264x352_b64"] = (264, 352, 64, 50000)
sb.CASES["q_260x346_b64"] = (260, 346, 64, 50000)
sb.CASES["q_264x346_b64"] = (264, 346, 64, 50000)
rows, pk = sb.run_cases(["a_260x352_b64", "a_264x352_b64", "q_260x346_b64", "q_264x346_b64"], {"warp"})
for r in rows:
    for k, v in r.items():
        if isinstance(v, dict):
            print(r["case"], k, round(v["us"], 1), "us", round(100 * v["frac_hbm"], 1), "%")

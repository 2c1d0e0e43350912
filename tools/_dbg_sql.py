import asyncio, json, sys
sys.path.insert(0, ".")
from mcp_context_forge_b200 import framework as fw
from mcp_context_forge_b200.plugins.sql_sanitizer import SQLSanitizerPlugin
from oracle import sql_sanitizer_ref as ref
gold = json.load(open("tests/golden/sql_sanitizer.json"))
loop = asyncio.new_event_loop()
nbad = 0
for block in gold:
    plug = SQLSanitizerPlugin(fw.PluginConfig(name="sql", kind="x", hooks=["tool_pre_invoke"], config=block["config"]))
    c = ref.config(block["config"])
    for case in block["cases"]:
        got = loop.run_until_complete(plug._scan_args(case["args"]))
        exp = ref.scan_args(case["args"], c)
        if got != exp and nbad < 4:
            nbad += 1
            print("CONFIG", block["config"]); print("ARGS", json.dumps(case["args"])); print("GOT", got); print("EXP", exp)
            slots = plug._slots(case["args"])
            texts = [s.text for s in slots]
            first = loop.run_until_complete(plug._batcher.scan_sub(plug._prog, texts, plug._rule_mask))
            for s, (b, new) in zip(slots, first):
                print("   ", repr(s.text), bin(b), new, "| oracle strip:", repr(ref.strip_comments(s.text)))
print("bad", nbad)

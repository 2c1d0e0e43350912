"""oracle/ — CPU restatement of the reference's algorithms for the plugin hook-chain hot path.

TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
`--impl reference` legs may import this package; the product (mcp_context_forge_b200/) never does.
Every function cites the reference file:line it restates (paths relative to /root/reference).
Pinned against golden vectors generated from the reference's own plugin files
(tests/golden/, tools/gen_golden.py) — see DESIGN.md "Oracle".
"""

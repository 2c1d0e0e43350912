"""`from mcp_context_forge_b200.framework import Plugin, ...` — the plugin framework surface the
GPU plugins subclass: the real `cpex.framework` when installed (drop-in deployment), otherwise the
in-tree restatement (cpex_compat)."""
from .cpex_compat import real_cpex_available

if real_cpex_available():  # pragma: no cover - depends on the deployment
    from cpex.framework import *  # type: ignore # noqa: F401,F403
    from cpex.framework import Plugin, PluginConfig, PluginContext, PluginViolation  # type: ignore # noqa: F401
else:
    from .cpex_compat.framework import *  # noqa: F401,F403

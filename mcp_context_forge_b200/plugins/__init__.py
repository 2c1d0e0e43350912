"""GPU drop-in plugins.  Swap the `kind:` lines of plugins/config.yaml:

  plugins.regex_filter.search_replace.SearchReplacePlugin
      -> mcp_context_forge_b200.plugins.regex_filter.SearchReplacePlugin
  plugins.deny_filter.deny.DenyListPlugin
      -> mcp_context_forge_b200.plugins.deny_filter.DenyListPlugin
  plugins.harmful_content_detector.harmful_content_detector.HarmfulContentDetectorPlugin
      -> mcp_context_forge_b200.plugins.harmful_content_detector.HarmfulContentDetectorPlugin
  plugins.toon_encoder.toon_encoder.ToonEncoderPlugin
      -> mcp_context_forge_b200.plugins.toon_encoder.ToonEncoderPlugin
  plugins.sql_sanitizer.sql_sanitizer.SQLSanitizerPlugin
      -> mcp_context_forge_b200.plugins.sql_sanitizer.SQLSanitizerPlugin
  plugins.code_safety_linter.code_safety_linter.CodeSafetyLinterPlugin
      -> mcp_context_forge_b200.plugins.code_safety_linter.CodeSafetyLinterPlugin
  plugins.json_repair.json_repair.JSONRepairPlugin
      -> mcp_context_forge_b200.plugins.json_repair.JSONRepairPlugin
"""

"""Stand-in for the two imsim helpers the plugin imports (camera lookup, data directory) -- TEST INFRASTRUCTURE."""

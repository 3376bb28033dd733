"""Empty: terrain generation is one-time CPU set-up and out of scope."""

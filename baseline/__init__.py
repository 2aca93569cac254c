"""Reference arm of the benchmark: loader + installer for the UNMODIFIED reference (test / bench infrastructure)."""

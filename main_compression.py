"""Reference-compatible entry point: `python main_compression.py --test_dir ... --prior_path ...`
(flags as in the reference driver, main_compression.py:12-23)."""
from recombiner_b200.main_compression import compress, load_prior, main, parse_args  # noqa: F401

if __name__ == '__main__':
    main()

{ "targets": [ { "target_name": "nd4b", "sources": ["nd4b_napi.cc"], "defines": ["ND4B_HAVE_NODE_API"],
    "include_dirs": ["../../include"], "libraries": ["-L<(module_root_dir)/..", "-lnd4b", "-Wl,-rpath,<(module_root_dir)/.."] } ] }

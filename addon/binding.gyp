{
  "targets": [{
    "target_name": "amgpu_napi",
    "sources": ["amgpu_napi.cc"],
    "include_dirs": ["../include"],
    "libraries": ["-L<(module_root_dir)/../automerge_classic_b200", "-lamgpu", "-Wl,-rpath,<(module_root_dir)/../automerge_classic_b200"],
    "cflags_cc": ["-std=c++17", "-O2"]
  }]
}

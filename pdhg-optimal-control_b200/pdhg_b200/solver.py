"""Pickle save / load of solutions — same on-disk layout as the reference (jaxsrc/solver.py:13-33):
final file = `(results, errs_all)`, middle file = `[max_iters, phi_all, rho_all, alp_all, errs_all]`."""
import os
import pickle


def save(save_dir, filename, results):
  if not os.path.exists(save_dir):
    os.makedirs(save_dir)
  filename_full = save_dir + '/{}.pickle'.format(filename)
  with open(filename_full, 'wb') as file:
    pickle.dump(results, file)
    print('saved to {}'.format(file), flush=True)


def load_solution(dir, filename):
  filename = dir + '/{}.pickle'.format(filename)
  with open(filename, 'rb') as f:
    results, errors = pickle.load(f)
    print('loaded from {}'.format(filename), flush=True)
  return results, errors


def load_middle_solution(dir, filename):
  filename = dir + '/{}.pickle'.format(filename)
  with open(filename, 'rb') as f:
    results = pickle.load(f)
    print('loaded from {}'.format(filename), flush=True)
  return results

"""Knowledge Tracing Machines driver -- drop-in for the reference's fm.py (same command line, same outputs:
<data>/<dataset>/<legend>/<run>/results.json with ACC / AUC / NLL, X.npz, vectors-<d>.npy), with the factorization
machine trained on the B200 by tf-recomm_b200's FM step instead of libFM's MCMC sampler.

Reference flow kept (fm.py:16-181): parse flags -> dataset paths + config.yml -> all.csv, q-matrix, per-skill
counters -> legend / experiment folders -> df_to_sparse (KTM encoding) -> 5 folds BY USER -> train, predict the
held-out users -> metrics -> results.json.  Replaced: `pywFM.FM(task='classification', learning_method='mcmc',
k2=d).run(...)` (fm.py:104-110,154-155: libFM binary, absent here and out of scope as an algorithm) by mini-batch
training of the same model (w0, W, V of forward.py:21-22) with sigmoid cross-entropy + L2 + TF-semantics Adam, the
step BASELINE.json's north_star names.  `--d 0` keeps the reference's sklearn LogisticRegression (fm.py:139-152).

Extra flags (all optional): --data_folder, --lr, --reg, --batch, --seed, --folds, --synthetic N (write a synthetic
ASSISTments-shaped dataset of N events under --dataset first; there is no network for the real one).
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def build_parser():
    parser = argparse.ArgumentParser(description='Run Knowledge Tracing Machines')
    parser.add_argument('--base_dir', type=str, nargs='?', default=os.getcwd())  # kept; only used for LIBFM_PATH upstream
    parser.add_argument('--libfm', type=str, nargs='?', default='code/libfm')    # kept for CLI compatibility, unused
    parser.add_argument('--dataset', type=str, nargs='?', default='dummy')
    parser.add_argument('--d', type=int, nargs='?', default=20)
    for agent in ('users', 'items', 'skills', 'attempts', 'wins', 'fails', 'item_wins', 'item_fails', 'extra'):
        parser.add_argument('--' + agent, type=bool, nargs='?', const=True, default=False)
    parser.add_argument('--iter', type=int, nargs='?', default=500)  # upstream: MCMC iterations; here: epochs
    parser.add_argument('--data_folder', type=str, default='data')
    parser.add_argument('--lr', type=float, default=1e-2)
    parser.add_argument('--reg', type=float, default=3e-2)
    parser.add_argument('--batch', type=int, default=0, help='rows per step; 0 = BATCH_SIZE of config.yml')
    parser.add_argument('--seed', type=int, default=None, help='fold shuffle + init seed (upstream: unseeded)')
    parser.add_argument('--folds', type=int, default=5)
    parser.add_argument('--synthetic', type=int, default=0)
    return parser


def train_fm_on_device(X_train, y_train, X_test, d, n_iter, lr, reg, batch, seed):
    """Replaces fm.py:154-155.  -> (probabilities on X_test, pairwise_interactions V, per-epoch log)."""
    import torch
    from tf_recomm_b200._lib import LOSS_SIGMOID_CE
    from tf_recomm_b200.fm_engine import FmEngine
    n, F = X_train.shape
    eng = FmEngine(F, d, lr, reg, flags=LOSS_SIGMOID_CE, seed=13575 if seed is None else seed)
    n_chunks = max(1, int(np.ceil(n / batch)))
    chunks = np.array_split(np.arange(n), n_chunks)  # OneEpochIterator's chunking (dataio.py:126)
    batches = [eng.upload_csr(X_train[c], y_train[c]) for c in chunks]
    test = eng.upload_csr(X_test)
    rlog = []
    for epoch in range(n_iter):
        eng.run_epoch(batches)
        if epoch % max(1, n_iter // 10) == 0 or epoch == n_iter - 1:
            rlog.append(dict(epoch=epoch, global_step=eng.global_step))
    logits = eng.forward(test)
    proba = torch.sigmoid(logits.double()).cpu().numpy()
    return proba, eng.get_tables()["V"], rlog


def main(argv=None):
    from scipy.sparse import save_npz
    from sklearn.metrics import accuracy_score, log_loss, roc_auc_score
    from sklearn.model_selection import KFold
    import dataio
    from tf_recomm_b200 import ktm
    options = build_parser().parse_args(argv)
    experiment_args = vars(options)
    DATASET_NAME = options.dataset
    if options.synthetic:
        shape = dict(n_events=options.synthetic)
        if options.synthetic < 100000:  # small smoke datasets: shrink the id spaces with the event count
            shape.update(user_num=max(10, options.synthetic // 80), item_num=max(10, options.synthetic // 13),
                         n_skills=max(3, min(123, options.synthetic // 300)))
        users, items, outcomes, q = ktm.make_ktm_events(seed=13575, **shape)
        ktm.write_dataset(DATASET_NAME, users, items, outcomes, q, int(users.max()) + 1, q.shape[0],
                          data_folder=options.data_folder)
    CSV_FOLDER = dataio.build_new_paths(DATASET_NAME, options.data_folder)[0]
    df, config, qmatrix, skill_wins, skill_fails = ktm.load_dataset(DATASET_NAME, options.data_folder)
    USER_NUM, ITEM_NUM = config['USER_NUM'], config['ITEM_NUM']
    short_legend, full_legend, latex_legend, active_agents = dataio.get_legend(experiment_args)
    EXPERIMENT_FOLDER = os.path.join(CSV_FOLDER, short_legend)
    dataio.prepare_folder(EXPERIMENT_FOLDER)
    for run_id in range(options.folds):
        dataio.prepare_folder(os.path.join(EXPERIMENT_FOLDER, str(run_id)))

    print(df.head())
    X_fm = ktm.df_to_sparse(df, active_agents, USER_NUM, ITEM_NUM, qmatrix, skill_wins, skill_fails)
    save_npz(os.path.join(EXPERIMENT_FOLDER, 'X.npz'), X_fm)
    print('DF shape', df.shape)
    print('Xb shape', X_fm.shape)
    y_fm = np.array(df['outcome'])
    print('Encoding done')
    batch = options.batch or int(config.get('BATCH_SIZE', 10000))

    # Run experiments by separating students (fm.py:113-131)
    kf = KFold(n_splits=options.folds, shuffle=True, random_state=options.seed)
    all_users = df['user'].unique()
    results = []
    for run_id, (i_user_train, i_user_test) in enumerate(kf.split(all_users)):
        users_train = all_users[i_user_train]
        in_train = np.isin(np.asarray(df['user']), users_train)
        i_train, i_test = np.flatnonzero(in_train), np.flatnonzero(~in_train)
        X_train, y_train = X_fm[i_train], y_fm[i_train]
        X_train.data = np.nan_to_num(X_train.data)
        X_test, y_test = X_fm[i_test], y_fm[i_test]
        X_test.data = np.nan_to_num(X_test.data)
        start = time.time()
        if options.d == 0:  # fm.py:139-152
            from sklearn.linear_model import LogisticRegression
            model = LogisticRegression()
            model.fit(X_train, y_train)
            y_pred_test = model.predict_proba(X_test)[:, 1]
        else:
            y_pred_test, V, rlog = train_fm_on_device(X_train, y_train, X_test, options.d, options.iter, options.lr,
                                                      options.reg, batch, options.seed)
            np.save(os.path.join(EXPERIMENT_FOLDER, str(run_id), 'vectors-{:d}.npy'.format(options.d)), V)
            with open(os.path.join(EXPERIMENT_FOLDER, str(run_id), 'rlog.csv'), 'w') as f:
                f.write('epoch,global_step\n' + ''.join('%d,%d\n' % (r['epoch'], r['global_step']) for r in rlog))
        print('fit', time.time() - start)
        ACC = accuracy_score(y_test, np.round(y_pred_test))
        print('acc', ACC)
        AUC = roc_auc_score(y_test, y_pred_test)
        print('auc', AUC)
        NLL = log_loss(y_test, y_pred_test)
        with open(os.path.join(EXPERIMENT_FOLDER, str(run_id), 'results.json'), 'w') as f:
            f.write(json.dumps({
                'args': experiment_args,
                'legends': {'short': short_legend, 'full': full_legend, 'latex': latex_legend},
                'metrics': {'ACC': ACC, 'AUC': AUC, 'NLL': NLL}
            }, indent=4))
        results.append(dict(ACC=ACC, AUC=AUC, NLL=NLL))
    return results


if __name__ == '__main__':
    main()

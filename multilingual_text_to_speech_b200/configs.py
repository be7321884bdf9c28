"""Named model configurations of BASELINE.json (`configs`), as hyper-parameter overlays on the defaults.

The values restate the reference's params/*.json (only the keys that change tensor shapes or the graph);
SURVEY.md appendix B is the table they were checked against.
"""
from .params.params import Params as hp

_CSS10_CHARS = " abcdefghijklmnopqrstuvwxyzçèéßäöōǎǐíǒàáǔüèéìūòóùúāēěīâêôûñőűабвгдежзийклмнопрстуфхцчшщъыьэюяёάέήίαβγδεζηθικλμνξοπρςíστυφχψωόύώ"
_CSS_COMVOI_CHARS = " abcdefghijklmnopqrstuvwxyzçèéßäöōǎǐíǒàáǔüèéìūòóùúāēěīâêôûñőűабвгдежзийклмнопрстуфхцчшщъыьэюяё"

CONFIGS = {
    # cfg 1: default Params == LJ Speech monolingual (reference params/params.py)
    'ljspeech': dict(),
    # cfg 2: params/generated_training.json (CSS10, 10 languages, generated encoder)
    'generated_training': dict(
        batch_size=60, case_sensitive=False, characters=_CSS10_CHARS, dataset='css10', encoder_dimension=256,
        encoder_type='generated', generator_bottleneck_dim=8, generator_dim=20,
        languages=['german', 'french', 'hungarian', 'chinese', 'spanish', 'dutch', 'finnish', 'russian', 'japanese', 'greek'],
        language_embedding_dimension=32, multi_language=True, perfect_sampling=True, balanced_sampling=True),
    # cfg 3: params/shared_switching.json (vanilla encoder + speaker/language embeddings + adversarial classifier)
    'shared_switching': dict(
        batch_size=50, case_sensitive=False, characters=_CSS_COMVOI_CHARS, dataset='css_comvoi', encoder_dimension=256,
        encoder_type='simple', languages=['de', 'fr', 'zh', 'ru', 'nl'], language_embedding_dimension=4, multi_language=True,
        multi_speaker=True, reversal_classifier=True, reversal_classifier_dim=256, reversal_classifier_w=0.5,
        reversal_gradient_clipping=0.25, speaker_embedding_dimension=32, balanced_sampling=True),
    # cfg 4/5: params/generated_switching.json
    'generated_switching': dict(
        batch_size=50, case_sensitive=False, characters=_CSS_COMVOI_CHARS, dataset='css_comvoi', encoder_dimension=256,
        encoder_type='generated', generator_bottleneck_dim=4, generator_dim=10, languages=['de', 'fr', 'zh', 'ru', 'nl'],
        language_embedding_dimension=0, multi_language=True, multi_speaker=True, perfect_sampling=True,
        balanced_sampling=True, reversal_classifier=True, reversal_classifier_dim=256, reversal_classifier_w=0.125,
        reversal_gradient_clipping=0.25, speaker_embedding_dimension=32),
}


def apply(name, speakers=7, **extra):
    """Reset hp, overlay the named configuration and the fields train.py fills in at start-up (train.py:239-240)."""
    hp.reset()
    hp.load_state_dict(CONFIGS[name])
    hp.load_state_dict(extra)
    hp.language_number = len(hp.languages) if hp.multi_language else 0
    hp.speaker_number = speakers if hp.multi_speaker else 0
    return hp


def as_namespace():
    import types
    return types.SimpleNamespace(**hp.state_dict())

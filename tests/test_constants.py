"""CPU: dataset topology tables equal the reference plugin constants (when /root/reference is present)."""
import os

import pytest

from openpifpaf_b200 import constants

REF = '/root/reference/src/openpifpaf/plugins'


@pytest.mark.skipif(not os.path.isdir(REF), reason='/root/reference absent')
def test_skeletons_equal_reference():
    ns = {}
    exec(open(os.path.join(REF, 'coco/constants.py')).read().split('KINEMATIC')[0], ns)
    assert ns['COCO_PERSON_SKELETON'] == constants.COCO_PERSON_SKELETON
    ns = {}
    exec(open(os.path.join(REF, 'wholebody/constants.py')).read().split('body_kps')[0], ns)
    assert ns['WHOLEBODY_SKELETON'] == constants.wholebody_skeleton()
